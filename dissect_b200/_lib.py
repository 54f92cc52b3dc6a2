"""ctypes binding of libdgrm.so (include/dgrm.h).  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGRM_LIB") or os.path.join(_HERE, "libdgrm.so")   # DGRM_LIB: experiment builds

K_NUM = 8
K_NAMES = ("count", "expand", "syrk", "paircnt", "final", "fetch", "gather", "d2h")

OPERANDS = {"fp16": 0, "bf16": 1, "fp16x2": 2}
COUNT_METHODS = {"mma": 0, "popc": 1, "fp4": 2}
FETCH = {"A": 0, "S": 1, "N": 2}
TRANSPORT = {"fp64": 0, "fp32": 1, "narrow": 2}

# every symbol include/dgrm.h declares
SYMBOLS = (
    "dgrm_default_config", "dgrm_device_count", "dgrm_validate_bed_header", "dgrm_create", "dgrm_destroy", "dgrm_last_error", "dgrm_abi_version",
    "dgrm_set_individuals", "dgrm_begin", "dgrm_push_bed_host", "dgrm_push_dosage_host", "dgrm_push_bgen_blocks_host", "dgrm_push_bed_device", "dgrm_release_inputs",
    "dgrm_load_bed_file", "dgrm_finish", "dgrm_sync", "dgrm_num_snps", "dgrm_num_individuals",
    "dgrm_snp_stats", "dgrm_bad_snps", "dgrm_fetch_block_cyclic", "dgrm_fetch_block_cyclic_cols", "dgrm_bind_output", "dgrm_unbind_outputs", "dgrm_bind_tiles", "dgrm_low_overlap_pairs",
    "dgrm_set_output_subset", "dgrm_write_grm_dat", "dgrm_write_grm_dat_part", "dgrm_num_tiles", "dgrm_tile_coords", "dgrm_fetch_tiles", "dgrm_fetch_tiles_range",
    "dgrm_host_alloc", "dgrm_host_free", "dgrm_kernel_times",
    "dgrm_store_bed_file", "dgrm_store_bed_host", "dgrm_store_bed_device", "dgrm_last_d2h_bytes", "dgrm_store_clear", "dgrm_stored_snps", "dgrm_push_stored",
    "dgrm_fetch_standardised", "dgrm_save", "dgrm_set_view", "dgrm_ld_begin", "dgrm_ld_pairs",
)
SAVE = {"copy": 0, "add": 1, "drop": 2}
VIEW = {"current": 0, "saved": 1, "saved-current": 2, "saved+current": 3}


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("operand", C.c_int32),
        ("count_method", C.c_int32), ("flat_norm", C.c_int32), ("chunk_snps", C.c_int32),
        ("cta_group", C.c_int32), ("profile", C.c_int32), ("resident_chunks", C.c_int32), ("resident_mb", C.c_int32),
        ("reserve_sms", C.c_int32), ("a_transport", C.c_int32), ("n_transport", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


_lib = None


def load():
    """Load libdgrm.so; raise if the CUDA extension has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  dissect_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    lib.dgrm_default_config.argtypes = [P(Config)]
    lib.dgrm_create.argtypes = [P(Config), P(vp)]
    lib.dgrm_destroy.argtypes = [vp]
    lib.dgrm_destroy.restype = None
    lib.dgrm_last_error.argtypes = [vp]
    lib.dgrm_last_error.restype = C.c_char_p
    lib.dgrm_abi_version.argtypes = []
    lib.dgrm_device_count.argtypes = []
    lib.dgrm_validate_bed_header.argtypes = [vp, C.c_char_p]
    lib.dgrm_set_individuals.argtypes = [vp, i64, vp]
    lib.dgrm_begin.argtypes = [vp, i64]
    lib.dgrm_push_bed_host.argtypes = [vp, vp, i64, i64, vp]
    lib.dgrm_push_dosage_host.argtypes = [vp, vp, i64, vp, i64, i64]
    lib.dgrm_push_bgen_blocks_host.argtypes = [vp, vp, vp, i64, C.c_int, i64, vp]
    lib.dgrm_push_bed_device.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.dgrm_release_inputs.argtypes = [vp, vp]
    lib.dgrm_load_bed_file.argtypes = [vp, C.c_char_p, i64, vp, i64]
    lib.dgrm_finish.argtypes = [vp]
    lib.dgrm_sync.argtypes = [vp]
    lib.dgrm_num_snps.argtypes = [vp]
    lib.dgrm_num_snps.restype = i64
    lib.dgrm_num_individuals.argtypes = [vp]
    lib.dgrm_num_individuals.restype = i64
    lib.dgrm_snp_stats.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.dgrm_bad_snps.argtypes = [vp, vp, i64, P(i64)]
    lib.dgrm_fetch_block_cyclic.argtypes = [vp, i32, vp, i64, i32, i32, i32, i32, i32, i32, C.c_char]
    lib.dgrm_fetch_block_cyclic_cols.argtypes = [vp, i32, vp, i64, i32, i32, i32, i32, i32, i32, C.c_char, i64, i64]
    lib.dgrm_bind_output.argtypes = [vp, i32, vp, i64, i32, i32, i32, i32, i32, i32, C.c_char]
    lib.dgrm_unbind_outputs.argtypes = [vp]
    lib.dgrm_bind_tiles.argtypes = [vp, i32, vp]
    lib.dgrm_low_overlap_pairs.argtypes = [vp, dbl, vp, i64, P(i64)]
    lib.dgrm_set_output_subset.argtypes = [vp, vp, i64]
    lib.dgrm_write_grm_dat.argtypes = [vp, C.c_char_p]
    lib.dgrm_write_grm_dat_part.argtypes = [vp, C.c_char_p, i32]
    lib.dgrm_num_tiles.argtypes = [vp]
    lib.dgrm_num_tiles.restype = i64
    lib.dgrm_tile_coords.argtypes = [vp, vp]
    lib.dgrm_fetch_tiles.argtypes = [vp, i32, vp]
    lib.dgrm_fetch_tiles_range.argtypes = [vp, i32, vp, i64, i64]
    lib.dgrm_host_alloc.argtypes = [i64]
    lib.dgrm_host_alloc.restype = vp
    lib.dgrm_host_free.argtypes = [vp]
    lib.dgrm_host_free.restype = None
    lib.dgrm_kernel_times.argtypes = [vp, P(dbl * K_NUM), P(i64 * K_NUM)]
    lib.dgrm_store_bed_file.argtypes = [vp, C.c_char_p, i64]
    lib.dgrm_store_bed_host.argtypes = [vp, vp, i64, i64]
    lib.dgrm_store_bed_device.argtypes = [vp, vp, i64, i64, vp]
    lib.dgrm_last_d2h_bytes.argtypes = [vp]
    lib.dgrm_ld_begin.argtypes = [vp, i64, vp, i64]
    lib.dgrm_ld_pairs.argtypes = [vp, dbl, vp, i64, P(i64)]
    lib.dgrm_last_d2h_bytes.restype = i64
    lib.dgrm_store_clear.argtypes = [vp]
    lib.dgrm_stored_snps.argtypes = [vp]
    lib.dgrm_stored_snps.restype = i64
    lib.dgrm_push_stored.argtypes = [vp, i64, i64, vp]
    lib.dgrm_fetch_standardised.argtypes = [vp, vp, i64, i32, i32, i32, i32, i32, i32]
    lib.dgrm_save.argtypes = [vp, i32]
    lib.dgrm_set_view.argtypes = [vp, i32]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:
            fn.restype = i32
    _lib = lib
    return lib
