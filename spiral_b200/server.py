"""Host-side mirror of the reference's server entry points for the Spiral path.

`SpiralServer` keeps the database and the client's public parameters resident in HBM and answers
queries; it is a thin Python face over the tier-3 C-ABI (sb200_server_*), which is the same code a
C++ host (spiral_b200/csrc/host_mirror.cpp) drives.  Names follow the reference:
process_query_fast (src/spiral.cpp:1584) = first_dim + fold; runConversionImproved (:2040) =
expand_and_convert.
"""
import ctypes as C

import numpy as np

from .lib import SpiralParams, check, load_library

N = 2048
_P64 = C.POINTER(C.c_uint64)
_P16 = C.POINTER(C.c_uint16)


def _p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P64)


class SpiralServer:
    def __init__(self, params: SpiralParams, device=0, rank=0, world=1, _view_of=None):
        self.lib = load_library()
        self.params = params
        self.rank, self.world = rank, world
        self.dim0, self.num_per = 1 << params.nu1, 1 << params.nu2
        self.local_num_per = self.num_per // world
        h = C.c_void_p()
        if _view_of is None:
            check(self.lib.sb200_server_create(C.byref(h), C.byref(params), device, rank, world), self.lib)
        else:
            check(self.lib.sb200_server_create_view(C.byref(h), _view_of.h), self.lib)
        self.h = h
        self._parent = _view_of                    # keeps the database owner alive

    def view(self):
        """A server for another concurrent client over the same resident database."""
        return SpiralServer(self.params, rank=self.rank, world=self.world, _view_of=self)

    # ---- database -----------------------------------------------------------------------
    def load_db_items(self, pts_u16, item_begin=0):
        """pts_u16: (items, 4, 2048) uint16, this shard's items in j-major order (load_db, src/spiral.cpp:1028)."""
        pts_u16 = np.ascontiguousarray(pts_u16, dtype=np.uint16)
        check(self.lib.sb200_server_load_db_items(self.h, pts_u16.ctypes.data_as(_P16), item_begin, pts_u16.shape[0]), self.lib)

    def load_db_reference(self, B):
        """B: the reference's own database buffer (layout of src/spiral.cpp:1139-1153)."""
        check(self.lib.sb200_server_load_db_reference(self.h, _p64(B)), self.lib)

    def load_db_implicit(self, B_slices, working_set):
        """The reference's --random-data database: `working_set` z-slices in load_db's layout; the scan reads slice z mod working_set."""
        check(self.lib.sb200_server_load_db_implicit(self.h, _p64(B_slices), working_set), self.lib)

    def load_db_implicit_constant(self, value, working_set):
        check(self.lib.sb200_server_load_db_implicit_constant(self.h, value, working_set), self.lib)

    def load_db_records(self, records):
        """records: the WHOLE database as the flat record stream (uint8 array or a file path); load_db's `has_data` branch."""
        if isinstance(records, (str, bytes)):
            path = records if isinstance(records, bytes) else records.encode()
            check(self.lib.sb200_server_load_db_records_file(self.h, path), self.lib)
        else:
            records = np.ascontiguousarray(records, dtype=np.uint8)
            check(self.lib.sb200_server_load_db_records(self.h, records.ctypes.data, records.size), self.lib)

    def save_db(self, path):
        """Snapshot of the preprocessed shard (load_db's `has_file && !load` branch)."""
        check(self.lib.sb200_server_save_db(self.h, path.encode()), self.lib)

    def load_db_snapshot(self, path):
        check(self.lib.sb200_server_load_db_snapshot(self.h, path.encode()), self.lib)

    @property
    def record_stream_bytes(self):
        return self.lib.sb200_server_record_stream_bytes(self.h)

    def shard_items(self, pts):
        """Select + order this shard's items from the full item-major plaintext array."""
        idx = [j * self.num_per + ii for j in range(self.dim0) for ii in range(self.rank, self.num_per, self.world)]
        return pts[idx]

    # ---- public parameters --------------------------------------------------------------
    def set_public_params(self, W_exp_left, W_exp_right, W_conv, V_conv):
        check(self.lib.sb200_server_set_public_params(self.h, _p64(W_exp_left), _p64(W_exp_right), _p64(W_conv), _p64(V_conv)), self.lib)

    # ---- query answering -----------------------------------------------------------------
    def answer(self, query_cv, stream=None):
        """Host query (2x1 ref-NTT, 64 KiB) -> host response (3x2 raw, row 0 mod q', rows 1-2 mod 4p)."""
        resp = np.empty(6 * N, dtype=np.uint64)
        check(self.lib.sb200_server_answer(self.h, query_cv.ctypes.data, resp.ctypes.data, stream), self.lib)
        return resp

    def answer_wire(self, wire, stream=None):
        """Wire query (uint8 array, include/spiral_b200.h "wire formats") -> QPBITS-packed response (uint64 words)."""
        wire = np.ascontiguousarray(wire, dtype=np.uint8)
        out = np.empty(self.lib.sb200_server_packed_response_bytes(self.h) // 8, dtype=np.uint64)
        check(self.lib.sb200_server_answer_wire(self.h, wire.ctypes.data, wire.size, out.ctypes.data, stream), self.lib)
        return out

    def unpack_response(self, packed):
        """Client-side inverse of the response packing: packed words -> 3x2 raw response."""
        resp = np.empty(6 * N, dtype=np.uint64)
        check(self.lib.sb200_unpack_response(_p64(resp), _p64(np.ascontiguousarray(packed)), 2 * N, 4 * N, self.params.qp_bits, self.params.p_db), self.lib)
        return resp

    def upload_query_wire(self, wire, stream=None):
        wire = np.ascontiguousarray(wire, dtype=np.uint8)
        check(self.lib.sb200_server_upload_query_wire(self.h, wire.ctypes.data, wire.size, stream), self.lib)

    def upload_query_wire_ptr(self, host_ptr, nbytes, stream=None):
        check(self.lib.sb200_server_upload_query_wire(self.h, host_ptr, nbytes, stream), self.lib)

    def upload_query(self, query_cv, stream=None):
        check(self.lib.sb200_server_upload_query(self.h, query_cv.ctypes.data, stream), self.lib)

    def upload_query_ptr(self, host_ptr, stream=None):
        """host_ptr: address of a (preferably pinned) 64 KiB ref-NTT query ciphertext."""
        check(self.lib.sb200_server_upload_query(self.h, host_ptr, stream), self.lib)

    def expand_and_convert(self, stream=None):
        check(self.lib.sb200_server_expand_and_convert(self.h, stream), self.lib)

    def process(self, resp_ptr=None, stream=None, marks=None):
        """All server stages of the uploaded query in one call; marks: None or a (c_void_p * 4) of cudaEvent_t handles."""
        check(self.lib.sb200_server_process(self.h, resp_ptr, stream, marks), self.lib)

    def prepare(self, resp_ptr=None, stream=None):
        """Build the CUDA graphs process() replays, without running anything."""
        check(self.lib.sb200_server_prepare(self.h, resp_ptr, stream), self.lib)

    def first_dim(self, stream=None):
        check(self.lib.sb200_server_first_dim(self.h, stream), self.lib)

    def scan(self, stream=None):
        check(self.lib.sb200_server_scan(self.h, stream), self.lib)

    @staticmethod
    def scan_batched(servers, stream=None):
        """One database pass for 2 or 4 servers sharing a database (a parent and its views)."""
        arr = (C.c_void_p * len(servers))(*[s.h for s in servers])
        check(servers[0].lib.sb200_server_scan_batched(arr, len(servers), stream), servers[0].lib)

    def enable_tc(self, capacity=16):
        """Build the tensor-core (limb-tile) copy of the resident database for batches of up to `capacity` queries."""
        check(self.lib.sb200_server_enable_tc(self.h, capacity), self.lib)

    def tc_only(self):
        """Free the scan-layout copy: the limb-tile copy is then the only one and single queries scan on the tensor-core kernel."""
        check(self.lib.sb200_server_tc_only(self.h), self.lib)

    @staticmethod
    def scan_batched_tc(servers, stream=None):
        """One tcgen05 pass over the database for up to 16 servers sharing it (enable_tc on the owner first)."""
        arr = (C.c_void_p * len(servers))(*[s.h for s in servers])
        check(servers[0].lib.sb200_server_scan_batched_tc(arr, len(servers), stream), servers[0].lib)

    def lift(self, stream=None):
        check(self.lib.sb200_server_lift(self.h, stream), self.lib)

    def copy_partial(self, dst_ptr, stream=None):
        check(self.lib.sb200_server_copy_partial(self.h, dst_ptr, stream), self.lib)

    def load_db_random(self, seed=1):
        check(self.lib.sb200_server_load_db_random(self.h, seed), self.lib)

    def fold_local(self, stream=None):
        check(self.lib.sb200_server_fold_local(self.h, stream), self.lib)

    def partial_ct_ptr(self):
        return self.lib.sb200_server_partial_ct(self.h)

    def fold_tail(self, gathered_ptr, resp_ptr, stream=None):
        check(self.lib.sb200_server_fold_tail(self.h, gathered_ptr, resp_ptr, stream), self.lib)

    # ---- exchange over NVLink peer memory (sharded servers) ----------------------------------
    def xchg_export(self):
        buf = C.create_string_buffer(self.lib.sb200_server_xchg_handle_bytes())
        check(self.lib.sb200_server_xchg_export(self.h, buf), self.lib)
        return buf.raw

    def xchg_connect(self, handles):
        """handles: list of world byte strings (every rank's xchg_export()), rank order; one process per GPU."""
        blob = b"".join(handles)
        check(self.lib.sb200_server_xchg_connect(self.h, blob), self.lib)

    def xchg_connect_local(self, servers):
        arr = (C.c_void_p * len(servers))(*[s.h for s in servers])
        check(self.lib.sb200_server_xchg_connect_local(self.h, arr), self.lib)

    def exchange_and_tail(self, resp_ptr, stream=None):
        check(self.lib.sb200_server_exchange_and_tail(self.h, resp_ptr, stream), self.lib)

    def xchg_error(self, stream=None):
        return self.lib.sb200_server_xchg_error(self.h, stream)

    def download(self, dev_ptr, words, stream=None):
        out = np.empty(words, dtype=np.uint64)
        check(self.lib.sb200_server_download(self.h, out.ctypes.data, dev_ptr, words, stream), self.lib)
        return out

    def first_dim_cts(self, stream=None):
        return self.download(self.lib.sb200_server_first_dim_cts(self.h), self.local_num_per * 6 * N, stream)

    @property
    def db_bytes(self):
        """ALGORITHMIC bytes of one scan of this shard (an implicit database holds fewer slices but every scan covers all 2048)."""
        return self.dim0 * self.local_num_per * 4 * N * 8

    def close(self):
        if self.h:
            self.lib.sb200_server_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PackServer:
    """SpiralPack / SpiralStreamPack resident server (testHighRate's server statements, src/testing.cpp:1007-1081):
    out_n^2 database planes of 1x1 plaintexts, 2x1 Regev ciphertexts, one packed (out_n+1) x out_n response.
    Sharded like SpiralServer: rank g owns the second-dimension indices ii = g (mod world) of every plane."""

    def __init__(self, params: SpiralParams, device=0, rank=0, world=1, _view_of=None, shard="nu2"):
        """shard: "nu2" = second dimension strided over the ranks (every plane on every rank); "planes" = whole planes
        p = rank (mod world) per rank (no exchange before the packing)."""
        self.lib = load_library()
        self.params = params
        self.rank, self.world = rank, world
        self.shard = shard if world > 1 else "nu2"
        self.dim0, self.num_per = 1 << params.nu1, 1 << params.nu2
        self.local_num_per = self.num_per if self.shard == "planes" else self.num_per // world
        self.planes = params.out_n * params.out_n
        h = C.c_void_p()
        if _view_of is None and self.shard == "planes":
            check(self.lib.sb200_pack_server_create_plane_sharded(C.byref(h), C.byref(params), device, rank, world), self.lib)
        elif _view_of is None:
            check(self.lib.sb200_pack_server_create_sharded(C.byref(h), C.byref(params), device, rank, world), self.lib)
        else:
            check(self.lib.sb200_pack_server_create_view(C.byref(h), _view_of.h), self.lib)
        self.h = h
        self._parent = _view_of                    # keeps the database owner alive

    def view(self):
        """A second client context (own keys, scratch) over this server's resident planes."""
        return PackServer(self.params, rank=self.rank, world=self.world, _view_of=self, shard=self.shard)

    def owns_plane(self, plane):
        return bool(self.lib.sb200_pack_server_owns_plane(self.h, plane))

    # ---- exchange over NVLink peer memory (sharded servers) ----------------------------------
    def xchg_export(self):
        buf = C.create_string_buffer(self.lib.sb200_pack_server_xchg_handle_bytes())
        check(self.lib.sb200_pack_server_xchg_export(self.h, buf), self.lib)
        return buf.raw

    def xchg_connect(self, handles):
        check(self.lib.sb200_pack_server_xchg_connect(self.h, b"".join(handles)), self.lib)

    def xchg_connect_local(self, servers):
        arr = (C.c_void_p * len(servers))(*[s.h for s in servers])
        check(self.lib.sb200_pack_server_xchg_connect_local(self.h, arr), self.lib)

    def exchange_and_tail(self, resp_ptr=None, stream=None):
        check(self.lib.sb200_pack_server_exchange_and_tail(self.h, resp_ptr, stream), self.lib)

    def xchg_error(self, stream=None):
        return self.lib.sb200_pack_server_xchg_error(self.h, stream)

    def upload_direct_split_ptr(self, firstdim_slice_ptr, folding_ptr, stream=None):
        check(self.lib.sb200_pack_server_upload_direct_split(self.h, firstdim_slice_ptr, folding_ptr, stream), self.lib)

    def process(self, resp_ptr=None, stream=None, marks=None):
        """All server stages of the uploaded query in one call (sharded servers: every rank calls it)."""
        check(self.lib.sb200_pack_server_process(self.h, resp_ptr, stream, marks), self.lib)

    def prepare(self, resp_ptr=None, stream=None):
        """Build the CUDA graphs process() replays, without running anything."""
        check(self.lib.sb200_pack_server_prepare(self.h, resp_ptr, stream), self.lib)

    def expansion_sharded(self):
        return bool(self.lib.sb200_pack_server_expansion_sharded(self.h))

    def enable_tc(self, capacity=16):
        check(self.lib.sb200_pack_server_enable_tc(self.h, capacity), self.lib)

    def tc_only(self):
        """Free the scan-layout planes (see SpiralServer.tc_only)."""
        check(self.lib.sb200_pack_server_tc_only(self.h), self.lib)

    @staticmethod
    def scan_batched_tc(servers, stream=None):
        """One tcgen05 pass over all planes for up to 16 pack servers sharing them (enable_tc on the owner first)."""
        arr = (C.c_void_p * len(servers))(*[s.h for s in servers])
        check(servers[0].lib.sb200_pack_server_scan_batched_tc(arr, len(servers), stream), servers[0].lib)

    # ---- database -----------------------------------------------------------------------
    def shard_items(self, plane_pts):
        """Select + order this shard's items from one plane's item-major (item = j*num_per + ii) plaintext array."""
        idx = [j * self.num_per + ii for j in range(self.dim0) for ii in range(self.rank, self.num_per, self.world)]
        return plane_pts[idx]

    def load_plane_items(self, plane, pts_u16):
        pts_u16 = np.ascontiguousarray(pts_u16, dtype=np.uint16)
        assert pts_u16.size == self.dim0 * self.local_num_per * N
        check(self.lib.sb200_pack_server_load_plane_items(self.h, plane, pts_u16.ctypes.data_as(_P16)), self.lib)

    def set_plane_item(self, plane, j, ii_local, poly_u16):
        """Replace one item of a loaded plane (first-dimension index j, shard-local second-dimension index ii_local)."""
        poly_u16 = np.ascontiguousarray(poly_u16, dtype=np.uint16)
        assert poly_u16.size == N
        check(self.lib.sb200_pack_server_set_plane_item(self.h, plane, j, ii_local, poly_u16.ctypes.data_as(_P16)), self.lib)

    def load_plane_reference(self, plane, db_buf):
        """db_buf: the WHOLE plane in the reference's convertDb layout (src/testing.cpp:316-340)."""
        check(self.lib.sb200_pack_server_load_plane_reference(self.h, plane, _p64(db_buf)), self.lib)

    def load_random(self, seed=1):
        check(self.lib.sb200_pack_server_load_random(self.h, seed), self.lib)

    def load_db_records(self, records):
        """records: the WHOLE database (all planes interleaved per item) as a uint8 array or a file path."""
        if isinstance(records, (str, bytes)):
            path = records if isinstance(records, bytes) else records.encode()
            check(self.lib.sb200_pack_server_load_db_records_file(self.h, path), self.lib)
        else:
            records = np.ascontiguousarray(records, dtype=np.uint8)
            check(self.lib.sb200_pack_server_load_db_records(self.h, records.ctypes.data, records.size), self.lib)

    def save_db(self, path):
        check(self.lib.sb200_pack_server_save_db(self.h, path.encode()), self.lib)

    def load_db_snapshot(self, path):
        check(self.lib.sb200_pack_server_load_db_snapshot(self.h, path.encode()), self.lib)

    def upload_query_wire(self, wire, stream=None):
        wire = np.ascontiguousarray(wire, dtype=np.uint8)
        check(self.lib.sb200_pack_server_upload_query_wire(self.h, wire.ctypes.data, wire.size, stream), self.lib)

    def answer_wire(self, wire, stream=None):
        wire = np.ascontiguousarray(wire, dtype=np.uint8)
        resp = np.empty(self.response_words, dtype=np.uint64)
        check(self.lib.sb200_pack_server_answer_wire(self.h, wire.ctypes.data, wire.size, resp.ctypes.data, None, stream), self.lib)
        return resp

    def db_words(self):
        """The scan-layout planes of this shard, read back from HBM (tests)."""
        return self.download(self.lib.sb200_pack_server_db_ptr(self.h), self.db_bytes // 8)

    def set_public_params(self, W_exp_left, W_exp_right, V, v_W):
        opt = lambda a: None if a is None else _p64(a)  # noqa: E731
        check(self.lib.sb200_pack_server_set_public_params(self.h, opt(W_exp_left), opt(W_exp_right), opt(V), _p64(v_W)), self.lib)

    # ---- query answering -----------------------------------------------------------------
    def answer(self, query_cv, want_cts=False, stream=None):
        resp = np.empty(self.response_words, dtype=np.uint64)
        cts = np.empty(self.planes * 2 * N, dtype=np.uint64) if want_cts else None
        check(self.lib.sb200_pack_server_answer(self.h, query_cv.ctypes.data, resp.ctypes.data, cts.ctypes.data if want_cts else None, stream), self.lib)
        return (resp, cts) if want_cts else resp

    def answer_direct(self, v_firstdim, v_folding, want_cts=False, stream=None):
        resp = np.empty(self.response_words, dtype=np.uint64)
        cts = np.empty(self.planes * 2 * N, dtype=np.uint64) if want_cts else None
        check(self.lib.sb200_pack_server_answer_direct(self.h, v_firstdim.ctypes.data, None if v_folding is None else v_folding.ctypes.data,
                                                       resp.ctypes.data, cts.ctypes.data if want_cts else None, stream), self.lib)
        return (resp, cts) if want_cts else resp

    def upload_query_ptr(self, host_ptr, stream=None):
        check(self.lib.sb200_pack_server_upload_query(self.h, host_ptr, stream), self.lib)

    def upload_direct_ptr(self, firstdim_ptr, folding_ptr, stream=None):
        check(self.lib.sb200_pack_server_upload_direct(self.h, firstdim_ptr, folding_ptr, stream), self.lib)

    def expand_and_convert(self, stream=None):
        check(self.lib.sb200_pack_server_expand_and_convert(self.h, stream), self.lib)

    def scan(self, stream=None):
        check(self.lib.sb200_pack_server_scan(self.h, stream), self.lib)

    def fold_local(self, stream=None):
        check(self.lib.sb200_pack_server_fold_local(self.h, stream), self.lib)

    def partial_cts_ptr(self):
        return self.lib.sb200_pack_server_partial_cts(self.h)

    @property
    def partial_words(self):
        return self.lib.sb200_pack_server_partial_words(self.h)

    def copy_partial(self, dst_ptr, stream=None):
        check(self.lib.sb200_pack_server_copy_partial(self.h, dst_ptr, stream), self.lib)

    def fold_tail(self, gathered_ptr, resp_ptr, stream=None):
        check(self.lib.sb200_pack_server_fold_tail(self.h, gathered_ptr, resp_ptr, stream), self.lib)

    def result_cts_ptr(self):
        return self.lib.sb200_pack_server_result_cts(self.h)

    def download(self, dev_ptr, words, stream=None):
        out = np.empty(words, dtype=np.uint64)
        check(self.lib.sb200_pack_server_download(self.h, out.ctypes.data, dev_ptr, words, stream), self.lib)
        return out

    @property
    def db_bytes(self):
        return self.lib.sb200_pack_server_db_bytes(self.h)

    @property
    def response_words(self):
        return self.lib.sb200_pack_server_response_words(self.h)

    def close(self):
        if self.h:
            self.lib.sb200_pack_server_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
