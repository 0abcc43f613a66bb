"""Host-side face of the GPU client (sb200_client_*): key generation, public parameters, query generation and decoding.

The reference's client lives in src/client.cpp and the client statements of src/spiral.cpp (runConversionImproved :2040-2335,
check_final :1428-1476); names follow it.  All randomness derives from the 32-byte seed."""
import ctypes as C

import numpy as np

from .lib import SpiralParams, check, load_library

N = 2048
_P64 = C.POINTER(C.c_uint64)


def _p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P64)


class SpiralClient:
    def __init__(self, params: SpiralParams, seed: bytes, device=0):
        assert len(seed) == 32
        self.lib = load_library()
        self.params = params
        h = C.c_void_p()
        self._seed = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
        check(self.lib.sb200_client_create(C.byref(h), C.byref(params), device, self._seed.ctypes.data), self.lib)
        self.h = h

    def public_params(self):
        """(W_exp_left, W_exp_right, W_conv, V_conv) as ref-NTT uint64 arrays - the arguments of SpiralServer.set_public_params."""
        polys = (C.c_size_t * 4)()
        check(self.lib.sb200_client_public_param_polys(self.h, polys), self.lib)
        mats = [np.zeros(int(n) * 2 * N, dtype=np.uint64) for n in polys]
        check(self.lib.sb200_client_public_params(self.h, *[_p64(m) for m in mats]), self.lib)
        return mats

    def query_wire(self, idx, query_id, wire_seed: bytes = None):
        """SEEDED wire query.  query_id must never repeat under one client seed; wire_seed=None (recommended) derives the
        row-0 seed from the client key and query_id, an explicit one must be fresh per query as well."""
        wire = np.zeros(self.lib.sb200_wire_query_bytes(1), dtype=np.uint8)
        if wire_seed is None:
            check(self.lib.sb200_client_query_wire(self.h, idx, query_id, None, wire.ctypes.data), self.lib)
            return wire
        assert len(wire_seed) == 32
        seed = np.frombuffer(bytes(wire_seed), dtype=np.uint8).copy()
        check(self.lib.sb200_client_query_wire(self.h, idx, query_id, seed.ctypes.data, wire.ctypes.data), self.lib)
        return wire

    def wire_seed(self, query_id):
        out = np.zeros(32, dtype=np.uint8)
        check(self.lib.sb200_client_wire_seed(self.h, query_id, out.ctypes.data), self.lib)
        return out.tobytes()

    def decode(self, total_resp):
        """3x2 raw response -> (4, 2048) plaintext coefficients (the record's 2x2 matrix of polynomials)."""
        pt = np.zeros(4 * N, dtype=np.uint64)
        check(self.lib.sb200_client_decode(self.h, _p64(np.ascontiguousarray(total_resp, dtype=np.uint64)), _p64(pt)), self.lib)
        return pt.reshape(4, N)

    def secret(self):
        sr, Sp = np.zeros(N, dtype=np.uint64), np.zeros(2 * N, dtype=np.uint64)
        check(self.lib.sb200_client_secret(self.h, _p64(sr), _p64(Sp)), self.lib)
        return sr, Sp

    def close(self):
        if self.h:
            self.lib.sb200_client_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PackClient(SpiralClient):
    """SpiralPack / SpiralStreamPack client on the GPU (sb200_pack_client_*; testHighRate's client statements,
    src/testing.cpp:904-1005, 1086-1122): out_n rows of S', packing keys, packed or direct-upload queries, out_n x out_n decode."""

    def __init__(self, params: SpiralParams, seed: bytes, device=0):
        assert len(seed) == 32
        self.lib = load_library()
        self.params = params
        h = C.c_void_p()
        self._seed = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
        check(self.lib.sb200_pack_client_create(C.byref(h), C.byref(params), device, self._seed.ctypes.data), self.lib)
        self.h = h
        self.planes = params.out_n * params.out_n

    def public_params(self, direct=False):
        """(W_exp_left, W_exp_right, V, v_W) - the arguments of PackServer.set_public_params; direct: the first three are None."""
        polys = (C.c_size_t * 4)()
        check(self.lib.sb200_pack_client_public_param_polys(self.h, polys), self.lib)
        vW = np.zeros(int(polys[3]) * 2 * N, dtype=np.uint64)
        if direct:
            check(self.lib.sb200_pack_client_public_params(self.h, None, None, None, _p64(vW)), self.lib)
            return None, None, None, vW
        mats = [np.zeros(int(n) * 2 * N, dtype=np.uint64) for n in polys[:3]]
        check(self.lib.sb200_pack_client_public_params(self.h, *[_p64(m) for m in mats], _p64(vW)), self.lib)
        return mats[0], mats[1], mats[2], vW

    def query_direct(self, idx, query_id=0):
        p = self.params
        v_first = np.zeros((1 << p.nu1) * 2 * 2 * N, dtype=np.uint64)
        v_fold = np.zeros(max(p.nu2, 1) * 2 * 2 * p.t_gsw * 2 * N, dtype=np.uint64)
        check(self.lib.sb200_pack_client_query_direct(self.h, idx, query_id, _p64(v_first), _p64(v_fold)), self.lib)
        return v_first, v_fold

    def decode(self, total_resp):
        """(out_n+1) x out_n raw response -> (out_n^2, 2048) plaintext coefficients, plane-major."""
        pt = np.zeros(self.planes * N, dtype=np.uint64)
        check(self.lib.sb200_pack_client_decode(self.h, _p64(np.ascontiguousarray(total_resp, dtype=np.uint64)), _p64(pt)), self.lib)
        return pt.reshape(self.planes, N)

    def secret(self):
        sr, Sp = np.zeros(N, dtype=np.uint64), np.zeros(self.params.out_n * N, dtype=np.uint64)
        check(self.lib.sb200_client_secret(self.h, _p64(sr), _p64(Sp)), self.lib)
        return sr, Sp
