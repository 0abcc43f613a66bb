"""Loader for oracle/liboracle.so - the CPU restatement of the reference.  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, as the checker."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

N = 2048
P = 268369921
B = 249561089
Q = P * B

CONFIGS = {
    "cfg1": dict(t_gsw=8, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=20, out_n=2, p_db=256),
    "cfg3": dict(t_gsw=8, t_conv=4, t_exp=16, t_exp_right=56, qp_bits=20, out_n=4, p_db=256),
    "cfg4": dict(t_gsw=3, t_conv=56, t_exp=56, t_exp_right=56, qp_bits=27, out_n=5, p_db=65536),
    "cfg5": dict(t_gsw=9, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=21, out_n=2, p_db=256),
}
SO_CASE_MAX_IN = 4
KIND_RAW, KIND_NTT, KIND_PACKED = 0, 1, 2


class SoParams(C.Structure):
    _fields_ = [("nu1", C.c_uint32), ("nu2", C.c_uint32), ("t_gsw", C.c_uint32), ("t_conv", C.c_uint32),
                ("t_exp", C.c_uint32), ("t_exp_right", C.c_uint32), ("qp_bits", C.c_uint32),
                ("out_n", C.c_uint32), ("p_db", C.c_uint64)]


class SoCaseIO(C.Structure):
    _fields_ = [("inp", C.POINTER(C.c_uint64) * SO_CASE_MAX_IN), ("in_words", C.c_size_t * SO_CASE_MAX_IN),
                ("out", C.POINTER(C.c_uint64)), ("out_words", C.c_size_t), ("out_kind", C.c_int)]


class SoShape(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("npolys", "dim0", "num_per", "g", "stopround", "max_bits_right", "cur_dim")]


def make_params(cfg, nu1=0, nu2=0):
    d = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return SoParams(nu1, nu2, d["t_gsw"], d["t_conv"], d["t_exp"], d["t_exp_right"], d["qp_bits"], d["out_n"], d["p_db"])


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        build()
    lib = C.CDLL(LIB)
    u64p, sz = C.POINTER(C.c_uint64), C.c_size_t
    lib.so_case_count.restype = C.c_int
    lib.so_case_name.restype = C.c_char_p
    lib.so_case_name.argtypes = [C.c_int]
    lib.so_case_shape.argtypes = [C.c_int, C.POINTER(SoParams), C.POINTER(SoShape)]
    lib.so_case_make_inputs.argtypes = [C.c_int, C.POINTER(SoParams), C.c_uint64, C.POINTER(SoCaseIO)]
    lib.so_case_run_oracle.argtypes = [C.c_int, C.POINTER(SoParams), C.POINTER(SoCaseIO)]
    lib.so_case_digest.restype = C.c_uint64
    lib.so_case_digest.argtypes = [C.POINTER(SoCaseIO)]
    lib.so_digest_kind.restype = C.c_uint64
    lib.so_digest_kind.argtypes = [u64p, sz, C.c_int]
    lib.so_case_free.argtypes = [C.POINTER(SoCaseIO)]
    lib.so_tables.restype = u64p
    lib.so_fnv1a64.restype = C.c_uint64
    lib.so_fnv1a64.argtypes = [u64p, sz]
    lib.so_arb_qprime.restype = C.c_uint64
    lib.so_arb_qprime.argtypes = [C.c_uint32]
    lib.so_spiral_expansion_shape.argtypes = [C.POINTER(SoParams), C.POINTER(sz), C.POINTER(sz)]
    lib.so_spiral_answer.restype = C.c_int
    lib.so_spiral_answer.argtypes = [C.POINTER(SoParams)] + [u64p] * 9
    lib.so_load_db.argtypes = [u64p, u64p, C.c_uint32, C.c_uint32, C.c_uint64]
    lib.so_multiply_query_by_database.argtypes = [u64p, u64p, u64p, sz, sz]
    lib.so_reorient_ciphertexts.argtypes = [u64p, u64p, sz, sz]
    lib.so_ntt_inv_and_crt_lift.argtypes = [u64p, u64p, sz]
    lib.so_fold_one_further_dimension.argtypes = [sz, sz, u64p, u64p, u64p, C.c_uint32]
    lib.so_to_ntt.argtypes = [u64p, u64p, sz]
    lib.so_from_ntt.argtypes = [u64p, u64p, sz]
    lib.so_get_rescaled.argtypes = [u64p, u64p, sz, C.c_uint64, C.c_uint64]
    lib.so_pack_answer.restype = C.c_int
    lib.so_pack_answer.argtypes = [C.POINTER(SoParams), C.c_int] + [u64p] * 10
    lib.so_pack_expansion_shape.argtypes = [C.POINTER(SoParams), C.POINTER(sz), C.POINTER(sz)]
    lib.so_convert_db.argtypes = [u64p, u64p, sz, sz, sz]
    lib.so_fast_multiply_dim1.argtypes = [u64p, u64p, u64p, sz, sz]
    lib.so_encode_plaintext.argtypes = [u64p, u64p, sz, C.c_uint64]
    lib.so_modswitch_coeff.restype = C.c_uint64
    lib.so_modswitch_coeff.argtypes = [C.c_uint64, C.c_uint64]
    lib.so_modswitch_coeff_x87.restype = C.c_uint64
    lib.so_modswitch_coeff_x87.argtypes = [C.c_uint64, C.c_uint64]
    lib.so_modswitch.argtypes = [u64p, u64p, sz, C.c_uint32]
    lib.so_packed_words.restype = sz
    lib.so_packed_words.argtypes = [sz, C.c_uint32]
    lib.so_read_arbitrary_bits.restype = C.c_uint64
    lib.so_read_arbitrary_bits.argtypes = [u64p, sz, sz]
    lib.so_client_new.restype = C.c_void_p
    lib.so_client_new.argtypes = [C.POINTER(SoParams), C.c_uint64, C.c_int]
    lib.so_client_free.argtypes = [C.c_void_p]
    lib.so_client_w_exp_right_count.restype = sz
    lib.so_client_w_exp_right_count.argtypes = [C.POINTER(SoParams)]
    lib.so_client_spiral_pub_params.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p]
    lib.so_client_spiral_query.argtypes = [C.c_void_p, sz, u64p]
    lib.so_client_spiral_decode.argtypes = [C.c_void_p, u64p, u64p]
    # wire / on-disk formats (oracle/wire_format.h)
    u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    lib.so_chacha20_block.argtypes = [u32p, C.c_uint32, u32p, u32p]
    lib.so_wire_seeded_row0.argtypes = [u8p, u64p]
    lib.so_wire_query_bytes.restype = sz
    lib.so_wire_query_bytes.argtypes = [C.c_uint32]
    lib.so_wire_query_expand.restype = C.c_int
    lib.so_wire_query_expand.argtypes = [u8p, sz, u64p]
    lib.so_wire_query_pack_full.argtypes = [u64p, u8p]
    lib.so_client_spiral_query_wire.argtypes = [C.c_void_p, sz, C.c_uint32, u8p]
    lib.so_records_to_plaintexts.argtypes = [u64p, u8p, sz, C.c_uint64]
    lib.so_pack_client_new.restype = C.c_void_p
    lib.so_pack_client_new.argtypes = [C.POINTER(SoParams), C.c_uint64]
    lib.so_pack_client_new_chacha.restype = C.c_void_p
    lib.so_pack_client_new_chacha.argtypes = [C.POINTER(SoParams), u8p]
    lib.so_pack_client_chacha_query_wire.argtypes = [C.c_void_p, sz, C.c_uint32, u8p, u8p]
    lib.so_pack_client_pub_params.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p]
    lib.so_pack_client_query.argtypes = [C.c_void_p, sz, u64p]
    lib.so_pack_client_query_direct.argtypes = [C.c_void_p, sz, u64p, u64p]
    lib.so_pack_client_decode.argtypes = [C.c_void_p, u64p, u64p]
    lib.so_client_new_chacha.restype = C.c_void_p
    lib.so_client_new_chacha.argtypes = [C.POINTER(SoParams), u8p]
    lib.so_client_gaussian_thresholds.argtypes = [C.c_void_p, u64p]
    lib.so_client_secret.argtypes = [C.c_void_p, u64p, u64p]
    lib.so_client_secret_n.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
    lib.so_client_chacha_query_wire.argtypes = [C.c_void_p, sz, C.c_uint32, u8p, u8p]
    _lib = lib
    return lib


WIRE_SEEDED, WIRE_FULL = 1, 2


def ptr8(a):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def wire_expand(lib, wire):
    """Wire query (bytes as a uint8 array) -> 2x1 ref-NTT ciphertext, or None when the oracle rejects the buffer."""
    q = np.zeros(2 * 2 * N, dtype=np.uint64)
    rc = lib.so_wire_query_expand(ptr8(wire), wire.size, ptr(q))
    return q if rc == 0 else None


def ptr(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def golden(cfg):
    with open(os.path.join(GOLDEN_DIR, f"ref_digests_{cfg}.json")) as f:
        return json.load(f)


class Case:
    """One parity case: seeded inputs + the oracle's output, as numpy arrays."""

    def __init__(self, lib, case_id, prm, seed):
        self.lib, self.id, self.prm = lib, case_id, prm
        self.io = SoCaseIO()
        lib.so_case_make_inputs(case_id, C.byref(prm), seed, C.byref(self.io))
        lib.so_case_run_oracle(case_id, C.byref(prm), C.byref(self.io))
        self.name = lib.so_case_name(case_id).decode()
        self.shape = SoShape()
        lib.so_case_shape(case_id, C.byref(prm), C.byref(self.shape))
        self.inputs = [np.ctypeslib.as_array(self.io.inp[k], shape=(self.io.in_words[k],)).copy()
                       if self.io.in_words[k] else None for k in range(SO_CASE_MAX_IN)]
        self.out = np.ctypeslib.as_array(self.io.out, shape=(self.io.out_words,)).copy()
        self.kind = self.io.out_kind
        self.digest = lib.so_case_digest(C.byref(self.io))
        lib.so_case_free(C.byref(self.io))


def canon(a, kind):
    """Canonical form for comparison: raw exact; NTT / packed residues reduced modulo their prime."""
    a = np.asarray(a, dtype=np.uint64)
    if kind == KIND_RAW:
        return a
    if kind == KIND_NTT:
        v = a.reshape(-1, 2, N).copy()
        v[:, 0, :] %= np.uint64(P)
        v[:, 1, :] %= np.uint64(B)
        return v.reshape(-1)
    lo = (a & np.uint64(0xFFFFFFFF)) % np.uint64(P)
    hi = (a >> np.uint64(32)) % np.uint64(B)
    return lo | (hi << np.uint64(32))


class SpiralSession:
    """Test-side client + CPU reference pipeline for one parameter set (small sizes only)."""

    def __init__(self, lib, cfg, nu1, nu2, seed=1, nonoise=False, chacha_seed=None):
        """chacha_seed (32 bytes): the counter-based client the CUDA client is compared with (so_client_new_chacha)."""
        self.lib, self.prm = lib, make_params(cfg, nu1, nu2)
        p = self.prm
        g, stop = C.c_size_t(), C.c_size_t()
        lib.so_spiral_expansion_shape(C.byref(p), C.byref(g), C.byref(stop))
        self.g, self.stopround = g.value, stop.value
        self.dim0, self.num_per = 1 << nu1, 1 << nu2
        self.total_n = self.dim0 * self.num_per
        if chacha_seed is None:
            self.client = lib.so_client_new(C.byref(p), seed, int(nonoise))
        else:
            self.client = lib.so_client_new_chacha(C.byref(p), ptr8(np.frombuffer(bytes(chacha_seed), dtype=np.uint8).copy()))
        PL = 2 * N
        n_right = lib.so_client_w_exp_right_count(C.byref(p))
        self.W_left = np.zeros(self.g * 2 * p.t_exp * PL, dtype=np.uint64)
        self.W_right = np.zeros(n_right * 2 * p.t_exp_right * PL, dtype=np.uint64)
        self.W_conv = np.zeros(3 * 2 * p.t_conv * PL, dtype=np.uint64)
        self.V_conv = np.zeros(3 * 2 * p.t_conv * PL, dtype=np.uint64)
        lib.so_client_spiral_pub_params(self.client, ptr(self.W_left), ptr(self.W_right), ptr(self.W_conv), ptr(self.V_conv))
        rng = np.random.default_rng(seed + 1000)
        self.pts = rng.integers(0, p.p_db, size=(self.total_n, 4, N), dtype=np.uint64)   # item-major plaintext matrices

    def reference_db(self):
        Bbuf = np.zeros(self.total_n * 4 * N, dtype=np.uint64)
        self.lib.so_load_db(ptr(Bbuf), ptr(np.ascontiguousarray(self.pts.reshape(-1))), self.prm.nu1, self.prm.nu2, self.prm.p_db)
        return Bbuf

    def query(self, idx):
        q = np.zeros(2 * 2 * N, dtype=np.uint64)
        self.lib.so_client_spiral_query(self.client, idx, ptr(q))
        return q

    def query_wire(self, idx, kind=WIRE_SEEDED):
        wire = np.zeros(self.lib.so_wire_query_bytes(kind), dtype=np.uint8)
        self.lib.so_client_spiral_query_wire(self.client, idx, kind, ptr8(wire))
        return wire

    def chacha_query_wire(self, idx, query_id, wire_seed):
        wire = np.zeros(self.lib.so_wire_query_bytes(WIRE_SEEDED), dtype=np.uint8)
        self.lib.so_client_chacha_query_wire(self.client, idx, query_id, ptr8(np.frombuffer(bytes(wire_seed), dtype=np.uint8).copy()), ptr8(wire))
        return wire

    def secret(self):
        sr, Sp = np.zeros(N, dtype=np.uint64), np.zeros(2 * N, dtype=np.uint64)
        self.lib.so_client_secret(self.client, ptr(sr), ptr(Sp))
        return sr, Sp

    def records(self):
        """The planted plaintexts as the flat record stream of the DB file format (log2(p_db) bits per coefficient)."""
        bits = int(self.prm.p_db).bit_length() - 1
        assert bits in (8, 16)
        return np.ascontiguousarray(self.pts.astype(np.uint8 if bits == 8 else np.uint16).reshape(-1)).view(np.uint8)

    def oracle_answer(self, q, Bbuf):
        final_ct = np.zeros(6 * N, dtype=np.uint64)
        resp = np.zeros(6 * N, dtype=np.uint64)
        first = np.zeros(self.num_per * 6 * N, dtype=np.uint64)
        rc = self.lib.so_spiral_answer(C.byref(self.prm), ptr(q), ptr(self.W_left), ptr(self.W_right), ptr(self.W_conv),
                                       ptr(self.V_conv), ptr(Bbuf), ptr(final_ct), ptr(resp), ptr(first))
        assert rc == 0
        return resp, final_ct, first

    def decode(self, resp):
        out = np.zeros(4 * N, dtype=np.uint64)
        self.lib.so_client_spiral_decode(self.client, ptr(np.ascontiguousarray(resp)), ptr(out))
        return out.reshape(4, N)

    def close(self):
        self.lib.so_client_free(self.client)


class PackSession:
    """Test-side SpiralPack / SpiralStreamPack client + CPU reference pipeline (small sizes only): real keys, packing keys,
    queries and decoding, so the Pack servers are checked on real encryptions and on "decoded == planted"."""

    def __init__(self, lib, cfg, nu1, nu2, direct, seed=1, chacha_seed=None):
        """chacha_seed (32 bytes): counter-based randomness (so_pack_client_new_chacha), the statement for a CUDA Pack client."""
        self.lib, self.prm, self.direct = lib, make_params(cfg, nu1, nu2), direct
        p = self.prm
        g, stop = C.c_size_t(), C.c_size_t()
        lib.so_pack_expansion_shape(C.byref(p), C.byref(g), C.byref(stop))
        self.g, self.stopround = g.value, stop.value
        self.dim0, self.num_per, self.n = 1 << nu1, 1 << nu2, p.out_n
        self.total_n, self.planes = self.dim0 * self.num_per, p.out_n * p.out_n
        if chacha_seed is None:
            self.client = lib.so_pack_client_new(C.byref(p), seed)
        else:
            self.client = lib.so_pack_client_new_chacha(C.byref(p), ptr8(np.frombuffer(bytes(chacha_seed), dtype=np.uint8).copy()))
        PL = 2 * N
        self.v_W = np.zeros(self.n * (self.n + 1) * p.t_conv * PL, dtype=np.uint64)
        if direct:
            self.W_left = self.W_right = self.V = None
            lib.so_pack_client_pub_params(self.client, None, None, None, ptr(self.v_W))
        else:
            self.W_left = np.zeros(self.g * 2 * p.t_exp * PL, dtype=np.uint64)
            self.W_right = np.zeros((self.stopround + 1) * 2 * p.t_exp_right * PL, dtype=np.uint64)
            self.V = np.zeros(2 * 2 * p.t_conv * PL, dtype=np.uint64)
            lib.so_pack_client_pub_params(self.client, ptr(self.W_left), ptr(self.W_right), ptr(self.V), ptr(self.v_W))
        rng = np.random.default_rng(seed + 2000)
        self.pts = rng.integers(0, p.p_db, size=(self.planes, self.total_n, N), dtype=np.uint64)      # [plane][item][z]

    def reference_planes(self):
        """The out_n^2 planes in the reference's convertDb layout, concatenated (what so_pack_answer scans)."""
        items, PL = self.total_n, 2 * N
        db = np.zeros(self.planes * items * N, dtype=np.uint64)
        for pl in range(self.planes):
            enc = np.zeros(items * N, dtype=np.uint64)
            self.lib.so_encode_plaintext(ptr(enc), ptr(np.ascontiguousarray(self.pts[pl].reshape(-1))), items * N, self.prm.p_db)
            ntt = np.zeros(items * PL, dtype=np.uint64)
            self.lib.so_to_ntt(ptr(ntt), ptr(enc), items)
            self.lib.so_convert_db(ptr(db[pl * items * N:(pl + 1) * items * N]), ptr(ntt), items, self.dim0, self.num_per)
        return db

    def query(self, idx):
        """(query_cv, v_firstdim, v_folding): the packed ciphertext, or the direct-upload ciphertexts."""
        PL, ell = 2 * N, self.prm.t_gsw
        if self.direct:
            v_first = np.zeros(self.dim0 * 2 * PL, dtype=np.uint64)
            v_fold = np.zeros(max(self.prm.nu2, 1) * 2 * 2 * ell * PL, dtype=np.uint64)
            self.lib.so_pack_client_query_direct(self.client, idx, ptr(v_first), ptr(v_fold))
            return None, v_first, v_fold
        q = np.zeros(2 * PL, dtype=np.uint64)
        self.lib.so_pack_client_query(self.client, idx, ptr(q))
        return q, None, None

    def chacha_query_wire(self, idx, query_id, wire_seed):
        wire = np.zeros(self.lib.so_wire_query_bytes(WIRE_SEEDED), dtype=np.uint8)
        self.lib.so_pack_client_chacha_query_wire(self.client, idx, query_id, ptr8(np.frombuffer(bytes(wire_seed), dtype=np.uint8).copy()), ptr8(wire))
        return wire

    def oracle_answer(self, query, db):
        q, v_first, v_fold = query
        dummy = np.zeros(2 * 2 * N, dtype=np.uint64)
        n = self.n
        resp = np.zeros((n + 1) * n * N, dtype=np.uint64)
        cts = np.zeros(self.planes * 2 * N, dtype=np.uint64)
        opt = lambda a: ptr(a if a is not None else dummy)  # noqa: E731
        rc = self.lib.so_pack_answer(C.byref(self.prm), int(not self.direct), opt(q), opt(self.W_left), opt(self.W_right), opt(self.V),
                                     opt(v_first), opt(v_fold), ptr(self.v_W), ptr(db), ptr(resp), ptr(cts))
        assert rc == 0
        return resp, cts

    def decode(self, resp):
        out = np.zeros(self.planes * N, dtype=np.uint64)
        self.lib.so_pack_client_decode(self.client, ptr(np.ascontiguousarray(resp)), ptr(out))
        return out.reshape(self.planes, N)

    def planted(self, idx):
        return self.pts[:, idx, :]

    def close(self):
        self.lib.so_client_free(self.client)
