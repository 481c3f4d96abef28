"""ctypes binding of libb200orb.so (C ABI: include/b200orb.h).  No CPU fallback: if the library is
missing or no CUDA device is visible, every compute entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ORB_LIB") or os.path.join(_HERE, "libb200orb.so")   # B200ORB_LIB: A/B runs of two builds of the library
_lib = None

E_ARG, E_CUDA, E_STATE, E_RANGE = -1, -2, -3, -4


class B200OrbError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200OrbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  pyorbslam_b200 has no CPU fallback.")
    l = C.CDLL(LIB_PATH)
    vp, i32, f32, f64 = C.c_void_p, C.c_int, C.c_float, C.c_double
    l.b200orb_last_error.restype = C.c_char_p
    l.b200orb_kernel_launches.restype = C.c_longlong
    l.b200orb_extractor_create.argtypes = [i32, f32, i32, i32, i32, i32, C.POINTER(vp)]
    l.b200orb_extractor_destroy.argtypes = [vp]
    l.b200orb_extractor_destroy.restype = None
    l.b200orb_get_levels.argtypes = [vp]
    l.b200orb_get_scale_factor.argtypes = [vp]
    l.b200orb_get_scale_factor.restype = f32
    for f in ("b200orb_get_scale_factors", "b200orb_get_inverse_scale_factors", "b200orb_get_scale_sigma_squares",
              "b200orb_get_inverse_scale_sigma_squares", "b200orb_get_features_per_level"):
        getattr(l, f).argtypes = [vp, vp]
    l.b200orb_extract.argtypes = [vp, vp, i32, i32, C.POINTER(i32)]
    l.b200orb_get_results.argtypes = [vp, vp, vp]
    l.b200orb_max_keypoints.argtypes = [vp]
    l.b200orb_level_size.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    l.b200orb_get_pyramid_level.argtypes = [vp, i32, vp]
    l.b200orb_get_pyramid_all.argtypes = [vp, vp, C.c_longlong]
    l.b200orb_get_level_image.argtypes = [vp, i32, i32, vp]
    l.b200orb_get_level_candidates.argtypes = [vp, i32, i32, vp, C.POINTER(i32)]
    l.b200orb_stereo.argtypes = [vp, vp, f64, f32, vp, vp, vp]
    l.b200orb_stereo_ex.argtypes = [vp, vp, f64, f32, i32, vp, vp, vp, vp]
    l.b200orb_batch_set_stereo_flags.argtypes = [vp, i32]
    l.b200orb_batch_set_copy_only.argtypes = [vp, i32]
    l.b200orb_stereo_host.argtypes = [i32, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, f64, f32, vp, vp, vp]
    l.b200orb_batch_create.argtypes = [i32, f32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]
    l.b200orb_batch_destroy.argtypes = [vp]
    l.b200orb_batch_destroy.restype = None
    l.b200orb_batch_max_pairs.argtypes = [vp]
    l.b200orb_batch_kp_capacity.argtypes = [vp]
    l.b200orb_batch_workspace_bytes.argtypes = [vp]
    l.b200orb_batch_workspace_bytes.restype = C.c_longlong
    l.b200orb_batch_run_device.argtypes = [vp, vp, vp, i32, f64, f32, vp, vp, vp, vp, vp, vp, vp]
    l.b200orb_batch_run_host.argtypes = [vp, vp, vp, i32, f64, f32, vp, vp, vp, vp, vp, vp]
    l.b200orb_batch_run_host_shard.argtypes = [vp, vp, vp, i32, f64, f32, vp, vp, vp, vp, vp, vp, i32, i32]
    l.b200orb_batch_status_device.argtypes = [vp, i32, vp, vp]
    l.b200orb_batch_status_host.argtypes = [vp, vp, i32]
    l.b200orb_batch_candidate_count.argtypes = [vp, i32, C.POINTER(C.c_longlong)]
    l.b200orb_batch_profile.argtypes = [vp, i32, i32]
    l.b200orb_batch_stage_launches.argtypes = [vp, vp]
    l.b200orb_host_chunk_schedule.argtypes = [i32, i32, i32, vp, i32]
    l.b200orb_plan_cells.argtypes = [i32, f32, i32, i32, i32, i32, i32, vp, i32]
    l.b200orb_batch_profile_read.argtypes = [vp, vp, C.POINTER(i32), C.POINTER(C.c_longlong)]
    l.b200orb_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    l.b200orb_host_free.argtypes = [vp]
    _lib = l
    return l


def plan_cells(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, H, W):
    """FAST cell grid the engine plans for an H x W image: int32 array [cells][level, iniX, iniY, cw, ch, cand_ofs] (host logic)."""
    import numpy as np
    args = (int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST), int(H), int(W))
    n = lib().b200orb_plan_cells(*args, None, 0)
    if n < 0:
        check(n)
    out = np.zeros((n, 6), np.int32)
    lib().b200orb_plan_cells(*args, out.ctypes.data_as(C.c_void_p), n)
    return out


def chunk_schedule(max_pairs, lanes, n_pairs):
    """Chunk sizes b200orb_batch_run_host cuts a job into (pure host logic, no GPU needed)."""
    n = lib().b200orb_host_chunk_schedule(int(max_pairs), int(lanes), int(n_pairs), None, 0)
    if n < 0:
        check(n)
    buf = (C.c_int32 * n)()
    lib().b200orb_host_chunk_schedule(int(max_pairs), int(lanes), int(n_pairs), buf, n)
    return list(buf)


def check(rc):
    """C status -> Python exception, mirroring how pybind translates the reference's C++ exceptions
    (std::logic_error -> RuntimeError).  Range errors map to IndexError like the reference's Python."""
    if rc == 0:
        return
    msg = lib().b200orb_last_error().decode()
    if rc == E_ARG:
        raise ValueError(msg)
    if rc == E_RANGE:
        raise IndexError(msg)
    raise B200OrbError(msg)


def kernel_launches():
    return int(lib().b200orb_kernel_launches())


def device_count():
    return int(lib().b200orb_device_count())
