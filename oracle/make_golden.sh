#!/bin/bash
# Regenerates tests/golden/ref_digests_<cfg>.json from the UNMODIFIED reference.
# TEST INFRASTRUCTURE; only runs where /root/reference exists (the build container).
# Every case is evaluated by both ISA builds of the reference (AVX2 and, when the host has it,
# AVX-512) and by the oracle; the script fails if any of them disagree.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SEED="${SEED:-20220368}"
make -C "$HERE" oracle ref golden >/dev/null
mkdir -p "$HERE/../tests/golden"
for cfg in cfg1 cfg3 cfg4 cfg5; do
    isas="avx2"; grep -q avx512f /proc/cpuinfo && isas="avx2 avx512"
    for isa in $isas; do
        "$HERE/_ref/ref_golden_${cfg}_${isa}" "$cfg" "$SEED" "/tmp/ref_digests_${cfg}_${isa}.json"
    done
    if [ -f "/tmp/ref_digests_${cfg}_avx512.json" ]; then
        cmp "/tmp/ref_digests_${cfg}_avx2.json" "/tmp/ref_digests_${cfg}_avx512.json"
    fi
    cp "/tmp/ref_digests_${cfg}_avx2.json" "$HERE/../tests/golden/ref_digests_${cfg}.json"
done
echo "golden digests written to tests/golden/"
