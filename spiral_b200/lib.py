"""ctypes binding of libspiral_b200.so (C-ABI declared in include/spiral_b200.h)."""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspiral_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "spiral_b200.h")
CSRC = os.path.join(_HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class SB200Error(RuntimeError):
    pass


class SpiralParams(C.Structure):
    """Mirror of sb200_params (the reference's -D macros, include/values.h:78-93, at run time)."""
    _fields_ = [("nu1", C.c_uint32), ("nu2", C.c_uint32), ("t_gsw", C.c_uint32), ("t_conv", C.c_uint32),
                ("t_exp", C.c_uint32), ("t_exp_right", C.c_uint32), ("qp_bits", C.c_uint32),
                ("out_n", C.c_uint32), ("p_db", C.c_uint64)]


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER_PATH]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "spiral_b200.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


def declared_symbols():
    """Every sb200_* function declared in include/spiral_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[A-Za-z0-9_]+)\s*\(", text)))


_lib = None


def load_library():
    """Load libspiral_b200.so; fails loudly when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SB200Error(f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`; "
                         "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    u64p, u32p, u16p, vp, sz = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.c_void_p, C.c_size_t
    sig = {
        "sb200_init": (C.c_int, [C.c_int]),
        "sb200_last_error": (C.c_char_p, []),
        "sb200_abi_version": (C.c_int, []),
        "sb200_arb_qprime": (C.c_uint64, [C.c_uint32]),
        "sb200_launch_count": (C.c_uint64, []),
        "sb200_kernel_log": (sz, [C.c_char_p, sz]),
        "sb200_kernel_log_reset": (None, []),
        "sb200_trace_enable": (C.c_int, [C.c_uint32]),
        "sb200_trace_read": (sz, [vp, sz, C.c_int]),
        "sb200_db_words": (sz, [C.c_uint32, C.c_uint32]),
        "sb200_fold_scratch_words": (sz, [sz, C.c_uint32]),
        # tier 1 (device pointers as integers)
        "sb200_dev_ntt_from_ref": (C.c_int, [vp, vp, sz, vp]),
        "sb200_dev_ntt_to_ref": (C.c_int, [vp, vp, sz, vp]),
        "sb200_dev_to_ntt": (C.c_int, [vp, vp, sz, vp]),
        "sb200_dev_from_ntt": (C.c_int, [vp, vp, sz, vp]),
        "sb200_dev_multiply": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
        "sb200_dev_automorph": (C.c_int, [vp, vp, sz, C.c_uint32, vp]),
        "sb200_dev_gadget_ntt": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
        "sb200_dev_rescale": (C.c_int, [vp, vp, sz, C.c_uint64, C.c_uint64, vp]),
        "sb200_dev_modswitch": (C.c_int, [vp, vp, sz, C.c_uint32, vp]),
        "sb200_dev_bitpack": (C.c_int, [vp, vp, sz, C.c_uint32, vp]),
        "sb200_packed_words": (sz, [sz, C.c_uint32]),
        "sb200_packed_response_words": (sz, [sz, sz, C.c_uint32, C.c_uint64]),
        "sb200_dev_pack_response": (C.c_int, [vp, vp, sz, sz, C.c_uint32, C.c_uint64, vp]),
        "sb200_unpack_response": (C.c_int, [u64p, u64p, sz, sz, C.c_uint32, C.c_uint64]),
        "sb200_modswitch": (C.c_int, [u64p, u64p, C.c_uint32]),
        "sb200_server_answer_packed": (C.c_int, [vp, vp, vp, vp]),
        "sb200_server_packed_response_bytes": (sz, [vp]),
        "sb200_dev_db_build": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, sz, sz, vp]),
        "sb200_dev_db_from_reference": (C.c_int, [vp, vp, sz, sz, sz, sz, vp]),
        "sb200_dev_reorient_query": (C.c_int, [vp, vp, sz, vp]),
        "sb200_dev_first_dim": (C.c_int, [vp, vp, vp, sz, sz, vp]),
        "sb200_dev_fold_round": (C.c_int, [vp, sz, vp, vp, C.c_uint32, vp, vp]),
        # tier 2 (host pointers)
        "sb200_to_ntt": (C.c_int, [u64p, u64p, sz]),
        "sb200_from_ntt": (C.c_int, [u64p, u64p, sz]),
        "sb200_ntt_forward": (C.c_int, [u64p, sz]),
        "sb200_ntt_inverse": (C.c_int, [u64p, sz]),
        "sb200_multiply": (C.c_int, [u64p, u64p, u64p, C.c_int, C.c_int, C.c_int]),
        "sb200_automorph": (C.c_int, [u64p, u64p, sz, C.c_uint32]),
        "sb200_gadget_invert": (C.c_int, [u64p, u64p, C.c_int, C.c_int, C.c_int]),
        "sb200_getRescaled": (C.c_int, [u64p, u64p, sz, C.c_uint64, C.c_uint64]),
        "sb200_load_db": (C.c_int, [u64p, u64p, C.c_uint32, C.c_uint32, C.c_uint64]),
        "sb200_reorientCiphertexts": (C.c_int, [u64p, u64p, sz, sz]),
        "sb200_multiplyQueryByDatabase": (C.c_int, [u64p, u64p, u64p, sz, sz]),
        "sb200_nttInvAndCrtLiftCiphertexts": (C.c_int, [u64p, u64p, sz]),
        "sb200_split_and_crt": (C.c_int, [u64p, u64p, sz, C.c_uint32]),
        "sb200_foldOneFurtherDimension": (C.c_int, [sz, sz, u64p, u64p, u64p, C.c_uint32]),
        "sb200_expandImproved": (C.c_int, [u64p, sz, C.c_uint32, u64p, u64p, C.c_uint32, sz, sz]),
        "sb200_scalToMat": (C.c_int, [u64p, u64p, u64p, C.c_uint32]),
        "sb200_regevToGSW": (C.c_int, [u64p, u64p, C.c_uint32, C.c_uint32, u64p, u64p]),
        "sb200_convertDb": (C.c_int, [u64p, u64p, sz, sz, sz]),
        "sb200_reorientCiphertextsDim1": (C.c_int, [u64p, u64p, sz, sz, sz]),
        "sb200_fastMultiplyQueryByDatabaseDim1": (C.c_int, [u64p, u64p, u64p, sz, sz]),
        "sb200_foldCiphertextsDim1": (C.c_int, [u64p, sz, u64p, u64p, C.c_uint32]),
        "sb200_regevToSimpleGsw": (C.c_int, [u64p, u64p, sz, u64p, C.c_uint32, C.c_uint32, C.c_uint32, sz, sz]),
        "sb200_pack": (C.c_int, [u64p, C.c_uint32, C.c_uint32, u64p, u64p]),
        "sb200_resident_reorientCiphertexts": (C.c_int, [vp, vp, u64p, sz]),
        "sb200_resident_multiplyQueryByDatabase": (C.c_int, [vp, vp, vp]),
        "sb200_resident_nttInvAndCrtLiftCiphertexts": (C.c_int, [vp, vp, vp]),
        "sb200_resident_reorient_Q": (C.c_int, [vp, vp, u64p]),
        "sb200_resident_foldOneFurtherDimension": (C.c_int, [vp, sz, vp, vp, u64p]),
        # tier 3
        "sb200_pack_server_create": (C.c_int, [C.POINTER(vp), C.POINTER(SpiralParams), C.c_int]),
        "sb200_pack_server_create_sharded": (C.c_int, [C.POINTER(vp), C.POINTER(SpiralParams), C.c_int, C.c_int, C.c_int]),
        "sb200_pack_server_destroy": (None, [vp]),
        "sb200_pack_server_create_plane_sharded": (C.c_int, [C.POINTER(vp), C.POINTER(SpiralParams), C.c_int, C.c_int, C.c_int]),
        "sb200_pack_server_owns_plane": (C.c_int, [vp, sz]),
        "sb200_pack_server_local_planes": (sz, [vp]),
        "sb200_pack_server_xchg_handle_bytes": (sz, []),
        "sb200_pack_server_xchg_export": (C.c_int, [vp, vp]),
        "sb200_pack_server_xchg_connect": (C.c_int, [vp, vp]),
        "sb200_pack_server_xchg_connect_local": (C.c_int, [vp, C.POINTER(vp)]),
        "sb200_pack_server_exchange_and_tail": (C.c_int, [vp, vp, vp]),
        "sb200_pack_server_xchg_error": (C.c_int, [vp, vp]),
        "sb200_pack_server_upload_direct_split": (C.c_int, [vp, vp, vp, vp]),
        "sb200_pack_server_process": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
        "sb200_pack_server_prepare": (C.c_int, [vp, vp, vp]),
        "sb200_pack_server_expansion_sharded": (C.c_int, [vp]),
        "sb200_pack_server_upload_query": (C.c_int, [vp, vp, vp]),
        "sb200_pack_server_expand_and_convert": (C.c_int, [vp, vp]),
        "sb200_pack_server_upload_direct": (C.c_int, [vp, vp, vp, vp]),
        "sb200_pack_server_scan": (C.c_int, [vp, vp]),
        "sb200_pack_server_fold_local": (C.c_int, [vp, vp]),
        "sb200_pack_server_scan_plane_host": (C.c_int, [vp, sz, u64p, u64p]),
        "sb200_pack_server_partial_cts": (vp, [vp]),
        "sb200_pack_server_partial_words": (sz, [vp]),
        "sb200_pack_server_copy_partial": (C.c_int, [vp, vp, vp]),
        "sb200_pack_server_fold_tail": (C.c_int, [vp, vp, vp, vp]),
        "sb200_pack_server_result_cts": (vp, [vp]),
        "sb200_pack_server_response_ptr": (vp, [vp]),
        "sb200_pack_server_download": (C.c_int, [vp, vp, vp, sz, vp]),
        "sb200_pack_server_load_plane_items": (C.c_int, [vp, sz, u16p]),
        "sb200_pack_server_load_plane_reference": (C.c_int, [vp, sz, u64p]),
        "sb200_pack_server_set_plane_item": (C.c_int, [vp, sz, sz, sz, u16p]),
        "sb200_pack_server_load_random": (C.c_int, [vp, C.c_uint64]),
        "sb200_pack_server_set_public_params": (C.c_int, [vp, u64p, u64p, u64p, u64p]),
        "sb200_pack_server_answer": (C.c_int, [vp, vp, vp, vp, vp]),
        "sb200_pack_server_answer_direct": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "sb200_pack_server_db_bytes": (sz, [vp]),
        "sb200_pack_server_response_words": (sz, [vp]),
        "sb200_server_create": (C.c_int, [C.POINTER(vp), C.POINTER(SpiralParams), C.c_int, C.c_int, C.c_int]),
        "sb200_server_destroy": (None, [vp]),
        "sb200_server_create_view": (C.c_int, [C.POINTER(vp), vp]),
        "sb200_server_load_db_items": (C.c_int, [vp, u16p, sz, sz]),
        "sb200_server_load_db_reference": (C.c_int, [vp, u64p]),
        "sb200_server_db_ptr": (vp, [vp]),
        "sb200_server_load_db_implicit": (C.c_int, [vp, u64p, sz]),
        "sb200_server_load_db_implicit_constant": (C.c_int, [vp, C.c_uint64, sz]),
        "sb200_server_db_slices": (sz, [vp]),
        "sb200_server_set_public_params": (C.c_int, [vp, u64p, u64p, u64p, u64p]),
        "sb200_server_answer": (C.c_int, [vp, vp, vp, vp]),
        "sb200_server_upload_query": (C.c_int, [vp, vp, vp]),
        "sb200_server_process": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
        "sb200_server_prepare": (C.c_int, [vp, vp, vp]),
        "sb200_server_expand_and_convert": (C.c_int, [vp, vp]),
        "sb200_server_first_dim": (C.c_int, [vp, vp]),
        "sb200_server_fold_local": (C.c_int, [vp, vp]),
        "sb200_server_scan": (C.c_int, [vp, vp]),
        "sb200_server_lift": (C.c_int, [vp, vp]),
        "sb200_server_scan_batched": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
        "sb200_pack_server_create_view": (C.c_int, [C.POINTER(vp), vp]),
        "sb200_pack_server_enable_tc": (C.c_int, [vp, C.c_int]),
        "sb200_pack_server_tc_only": (C.c_int, [vp]),
        "sb200_pack_server_scan_batched_tc": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
        "sb200_fastMultiplyQueryByDatabaseDim1_batched": (C.c_int, [C.POINTER(u64p), u64p, C.POINTER(u64p), C.c_int, sz, sz]),
        "sb200_server_enable_tc": (C.c_int, [vp, C.c_int]),
        "sb200_server_tc_only": (C.c_int, [vp]),
        "sb200_server_scan_batched_tc": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
        "sb200_tc_supported": (C.c_int, [sz, sz]),
        "sb200_tc_query_bytes": (sz, [sz, C.c_int]),
        "sb200_dev_db_to_tc": (C.c_int, [vp, vp, sz, sz, vp]),
        "sb200_dev_query_to_tc": (C.c_int, [vp, vp, C.c_int, C.c_int, sz, vp]),
        "sb200_tc_scratch_bytes": (sz, [sz, C.c_int]),
        "sb200_dev_first_dim_tc": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, vp, vp, sz, sz, vp, vp]),
        "sb200_multiplyQueryByDatabase_batched": (C.c_int, [C.POINTER(u64p), C.POINTER(u64p), C.c_int, u64p, sz, sz]),
        "sb200_server_copy_partial": (C.c_int, [vp, vp, vp]),
        "sb200_server_scan_host": (C.c_int, [vp, u64p, u64p]),
        "sb200_server_load_db_random": (C.c_int, [vp, C.c_uint64]),
        "sb200_server_partial_ct": (vp, [vp]),
        "sb200_server_fold_tail": (C.c_int, [vp, vp, vp, vp]),
        "sb200_server_download": (C.c_int, [vp, vp, vp, sz, vp]),
        "sb200_server_xchg_handle_bytes": (sz, []),
        "sb200_server_xchg_export": (C.c_int, [vp, vp]),
        "sb200_server_xchg_connect": (C.c_int, [vp, vp]),
        "sb200_server_xchg_connect_local": (C.c_int, [vp, C.POINTER(vp)]),
        "sb200_server_exchange_and_tail": (C.c_int, [vp, vp, vp]),
        "sb200_server_xchg_error": (C.c_int, [vp, vp]),
        "sb200_server_xchg_reset": (C.c_int, [vp]),
        "sb200_server_expansion_sharded": (C.c_int, [vp]),
        "sb200_server_public_param_polys": (C.c_int, [vp, C.POINTER(sz)]),
        "sb200_client_wire_seed": (C.c_int, [vp, C.c_uint32, vp]),
        "sb200_server_first_dim_cts": (vp, [vp]),
        "sb200_server_query_bytes": (sz, [vp]),
        "sb200_server_response_bytes": (sz, [vp]),
        # client on the GPU
        "sb200_client_create": (C.c_int, [C.POINTER(vp), C.POINTER(SpiralParams), C.c_int, vp]),
        "sb200_client_destroy": (None, [vp]),
        "sb200_client_public_param_polys": (C.c_int, [vp, C.POINTER(sz)]),
        "sb200_client_public_params": (C.c_int, [vp, u64p, u64p, u64p, u64p]),
        "sb200_client_query_wire": (C.c_int, [vp, sz, C.c_uint32, vp, vp]),
        "sb200_client_decode": (C.c_int, [vp, u64p, u64p]),
        "sb200_client_secret": (C.c_int, [vp, u64p, u64p]),
        "sb200_client_gaussian_thresholds": (C.c_int, [u64p]),
        "sb200_pack_client_create": (C.c_int, [C.POINTER(vp), C.POINTER(SpiralParams), C.c_int, vp]),
        "sb200_pack_client_public_param_polys": (C.c_int, [vp, C.POINTER(sz)]),
        "sb200_pack_client_public_params": (C.c_int, [vp, u64p, u64p, u64p, u64p]),
        "sb200_pack_client_query_wire": (C.c_int, [vp, sz, C.c_uint32, vp, vp]),
        "sb200_pack_client_query_direct": (C.c_int, [vp, sz, C.c_uint32, u64p, u64p]),
        "sb200_pack_client_decode": (C.c_int, [vp, u64p, u64p]),
        # wire / on-disk formats
        "sb200_wire_query_bytes": (sz, [C.c_uint32]),
        "sb200_dev_query_from_wire": (C.c_int, [vp, vp, C.c_uint32, vp]),
        "sb200_server_upload_query_wire": (C.c_int, [vp, vp, sz, vp]),
        "sb200_server_answer_wire": (C.c_int, [vp, vp, sz, vp, vp]),
        "sb200_pack_server_upload_query_wire": (C.c_int, [vp, vp, sz, vp]),
        "sb200_pack_server_answer_wire": (C.c_int, [vp, vp, sz, vp, vp, vp]),
        "sb200_pack_server_db_ptr": (vp, [vp]),
        "sb200_server_record_stream_bytes": (sz, [vp]),
        "sb200_server_load_db_records": (C.c_int, [vp, vp, sz]),
        "sb200_server_load_db_records_file": (C.c_int, [vp, C.c_char_p]),
        "sb200_server_save_db": (C.c_int, [vp, C.c_char_p]),
        "sb200_server_load_db_snapshot": (C.c_int, [vp, C.c_char_p]),
        "sb200_pack_server_record_stream_bytes": (sz, [vp]),
        "sb200_pack_server_load_db_records": (C.c_int, [vp, vp, sz]),
        "sb200_pack_server_load_db_records_file": (C.c_int, [vp, C.c_char_p]),
        "sb200_pack_server_save_db": (C.c_int, [vp, C.c_char_p]),
        "sb200_pack_server_load_db_snapshot": (C.c_int, [vp, C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError here = header/library drift, caught by the CPU tests
        fn.restype, fn.argtypes = res, args
    lib._sb200_signatures = sig
    _lib = lib
    return lib


def kernel_log(lib=None):
    """Sorted list of the distinct kernel names launched since the last sb200_kernel_log_reset()."""
    lib = lib or load_library()
    n = lib.sb200_kernel_log(None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.sb200_kernel_log(buf, n + 1)
    return [k for k in buf.value.decode().split(";") if k]


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load_library()
        raise SB200Error(f"libspiral_b200 error {rc}: {lib.sb200_last_error().decode()}")
