"""GPU suite: the CUDA path, called through the C-ABI with the reference's host layouts, against the
oracle on the same seeded inputs (the 26 parity cases that are pinned to the reference itself).
Bar: bit-exact.  Raw-domain buffers are compared exactly; NTT-domain buffers are compared after
reducing both sides modulo the prime (the reference's own NTT may emit q for 0)."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu
P64 = C.POINTER(C.c_uint64)


def p(a):
    return a.ctypes.data_as(P64)


def ok(sb, rc):
    assert rc == 0, sb.sb200_last_error().decode()


def run_case(sb, case, prm):
    """Dispatch one oracle case to the C-ABI entry that replaces the same reference function."""
    s, i, N = case.shape, case.inputs, ol.N
    name = case.name
    out = np.zeros(case.out.size, dtype=np.uint64)
    if name in ("ntt_forward", "ntt_inverse"):
        out[:] = i[0]
        ok(sb, (sb.sb200_ntt_forward if name == "ntt_forward" else sb.sb200_ntt_inverse)(p(out), s.npolys))
    elif name in ("to_ntt", "to_ntt_no_reduce"):
        ok(sb, sb.sb200_to_ntt(p(out), p(i[0]), s.npolys))
    elif name == "from_ntt":
        ok(sb, sb.sb200_from_ntt(p(out), p(i[0]), s.npolys))
    elif name == "multiply":
        ok(sb, sb.sb200_multiply(p(out), p(i[0]), p(i[1]), 2, 3, 2))
    elif name == "automorph":
        ok(sb, sb.sb200_automorph(p(out), p(i[0]), s.npolys, N // 4 + 1))
    elif name == "gadget_invert":
        ok(sb, sb.sb200_gadget_invert(p(out), p(i[0]), 2 * prm.t_conv, 2, 1))
    elif name == "rescale":
        a, b = np.ascontiguousarray(i[0][:2 * N]), np.ascontiguousarray(i[0][2 * N:])
        oa, ob = np.zeros(2 * N, dtype=np.uint64), np.zeros(2 * N, dtype=np.uint64)
        ok(sb, sb.sb200_getRescaled(p(oa), p(a), 2 * N, ol.Q, sb.sb200_arb_qprime(prm.qp_bits)))
        ok(sb, sb.sb200_getRescaled(p(ob), p(b), 2 * N, ol.Q, 4 * prm.p_db))
        out = np.concatenate([oa, ob])
    elif name == "reorient_ciphertexts":
        ok(sb, sb.sb200_reorientCiphertexts(p(out), p(i[0]), s.dim0, 4))
    elif name == "first_dim":
        ok(sb, sb.sb200_multiplyQueryByDatabase(p(out), p(i[0]), p(i[1]), s.dim0, s.num_per))
    elif name == "ntt_inv_crt_lift":
        ok(sb, sb.sb200_nttInvAndCrtLiftCiphertexts(p(out), p(i[0]), s.num_per))
    elif name == "split_and_crt":
        ok(sb, sb.sb200_split_and_crt(p(out), p(i[0]), s.num_per, prm.t_gsw))
    elif name == "fold_one":
        cts = i[0].copy()
        ok(sb, sb.sb200_foldOneFurtherDimension(s.cur_dim, s.num_per, p(i[1]), p(i[2]), p(cts), prm.t_gsw))
        out = cts[: s.num_per * 6 * N]
    elif name in ("expand_full", "expand_stopround"):
        out[:] = i[0]
        ok(sb, sb.sb200_expandImproved(p(out), s.g, prm.t_exp, p(i[1]), p(i[2]), prm.t_exp_right, s.max_bits_right, s.stopround))
    elif name == "scal_to_mat":
        ok(sb, sb.sb200_scalToMat(p(out), p(i[0]), p(i[1]), prm.t_conv))
    elif name == "regev_to_gsw":
        ok(sb, sb.sb200_regevToGSW(p(out), p(i[0]), prm.t_conv, prm.t_gsw, p(i[1]), p(i[2])))
    elif name == "load_db":
        nu1, nu2 = int(s.dim0).bit_length() - 1, int(s.num_per).bit_length() - 1
        ok(sb, sb.sb200_load_db(p(out), p(i[0]), nu1, nu2, prm.p_db))
    elif name == "convert_db":
        ok(sb, sb.sb200_convertDb(p(out), p(i[0]), s.dim0 * s.num_per, s.dim0, s.num_per))
    elif name == "reorient_dim1":
        ok(sb, sb.sb200_reorientCiphertextsDim1(p(out), p(i[0]), 2 * s.dim0, s.dim0, 2))
    elif name == "first_dim_pack":
        ok(sb, sb.sb200_fastMultiplyQueryByDatabaseDim1(p(out), p(i[1]), p(i[0]), s.dim0, s.num_per))
    elif name == "fold_dim1":
        cts = i[0].copy()
        ok(sb, sb.sb200_foldCiphertextsDim1(p(cts), s.num_per, p(i[1]), p(i[2]), prm.t_gsw))
        out = cts[: 2 * N]
    elif name == "regev_to_simple_gsw":
        ok(sb, sb.sb200_regevToSimpleGsw(p(out), p(i[0]), 4 * prm.t_gsw + 2, p(i[1]), prm.t_conv, prm.t_gsw, 2, 2, 1))
    elif name == "pack":
        ok(sb, sb.sb200_pack(p(out), prm.out_n, prm.t_conv, p(i[0]), p(i[1])))
    elif name == "modswitch":
        ok(sb, sb.sb200_modswitch(p(out), p(i[0]), prm.qp_bits))
    else:
        return None
    return out


SPIRAL_CASES = ["ntt_forward", "ntt_inverse", "to_ntt", "to_ntt_no_reduce", "from_ntt", "multiply", "automorph",
                "gadget_invert", "rescale", "reorient_ciphertexts", "first_dim", "ntt_inv_crt_lift", "split_and_crt",
                "fold_one", "expand_full", "expand_stopround", "scal_to_mat", "regev_to_gsw", "load_db",
                "convert_db", "reorient_dim1", "first_dim_pack", "fold_dim1", "regev_to_simple_gsw", "pack", "modswitch"]


def test_every_golden_case_is_dispatched(oracle):
    assert sorted(SPIRAL_CASES) == sorted(ol.golden("cfg1")["cases"])


@pytest.mark.parametrize("cfg", ["cfg1", "cfg5", "cfg4", "cfg3"])
@pytest.mark.parametrize("name", SPIRAL_CASES)
def test_case_matches_oracle(sb, oracle, cfg, name):
    g = ol.golden(cfg)
    prm = ol.make_params(cfg)
    case = ol.Case(oracle, g["cases"][name]["id"], prm, g["seed"])
    assert f"{case.digest:016x}" == g["cases"][name]["digest"], "oracle drifted from the reference golden"
    got = run_case(sb, case, prm)
    assert got is not None
    want = ol.canon(case.out, case.kind)
    got = ol.canon(got, case.kind)
    assert got.shape == want.shape
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"{name}/{cfg}: {bad.size} of {want.size} words differ, first at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"
    # and the CUDA output hashes to the REFERENCE's own digest
    dig = oracle.so_digest_kind(p(np.ascontiguousarray(got)), got.size, case.kind)
    assert f"{dig:016x}" == g["cases"][name]["digest"]


@pytest.mark.parametrize("seed", [1, 7, 12345])
def test_other_seeds(sb, oracle, seed):
    prm = ol.make_params("cfg1")
    for name in ("from_ntt", "first_dim", "fold_one", "expand_stopround", "regev_to_gsw"):
        cid = ol.golden("cfg1")["cases"][name]["id"]
        case = ol.Case(oracle, cid, prm, seed)
        got = ol.canon(run_case(sb, case, prm), case.kind)
        assert np.array_equal(got, ol.canon(case.out, case.kind)), f"{name} seed {seed}"
