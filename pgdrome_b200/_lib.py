"""ctypes binding of libpgdb200.so (the C ABI in include/pgd_b200.h).

There is NO CPU fallback: every wrapper takes CUDA torch tensors and raises if the library is
missing or a tensor is not on the GPU.  PyTorch is used for device memory and streams only.
"""
import ctypes
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGD_LIB_PATH") or os.path.join(_HERE, "libpgdb200.so")  # override: kernel-variant experiments only

c_i32, c_i64, c_dbl, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p

# name -> argtypes (restype is int32 unless noted); mirrors include/pgd_b200.h one to one
SIGNATURES = {
    "pgd_abi_version": [],
    "pgd_create": [c_i32, ctypes.POINTER(c_vp)],
    "pgd_destroy": [c_vp],
    "pgd_last_error": [c_vp],
    "pgd_get_stats": [c_vp, c_vp, c_vp, c_i32],
    "pgd_set_option": [c_vp, ctypes.c_char_p, c_i64],
    "pgd_pattern_build_sync": [c_vp, c_vp, c_i64, c_i32, c_i64, ctypes.POINTER(c_i64), c_vp],
    "pgd_pattern_export": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "pgd_vecmap_build_sync": [c_vp, c_vp, c_i64, c_i32, c_i64, c_vp, c_vp, c_vp],
    "pgd_p1_rowplan_build_sync": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp],
    "pgd_assemble_p1_rows": [c_vp, c_vp, c_vp, c_i64, c_i32, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp],
    "pgd_assemble_p1_tensor": [c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp],
    "pgd_elem_bilinear": [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "pgd_elem_linear": [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "pgd_gather_values": [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp],
    "pgd_assemble_p1": [c_vp, c_vp, c_vp, c_i64, c_i32, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp],
    "pgd_lincomb": [c_vp, c_i32, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp],
    "pgd_apply_dirichlet": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp],
    "pgd_set_entries": [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp],
    "pgd_spmv": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp],
    "pgd_spmv_dot": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp],
    "pgd_bilinear": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp],
    "pgd_dot": [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp],
    "pgd_panel_dots": [c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp, c_vp],
    "pgd_pcg_sync": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_i32, c_i32, c_i32, c_i32, c_vp,
                     ctypes.POINTER(c_i32), ctypes.POINTER(c_dbl), c_vp],
    "pgd_spcg_init": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp],
    "pgd_spcg_init_fin": [c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp],
    "pgd_spcg_direction": [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp],
    "pgd_spcg_matvec": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp],
    "pgd_spcg_update": [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp],
    "pgd_spcg_rotate": [c_vp, c_vp, c_vp, c_vp],
    "pgd_comm_unique_id": [c_vp],
    "pgd_comm_init": [c_vp, c_vp, c_i32, c_i32],
    "pgd_comm_destroy": [c_vp],
    "pgd_spcg_solve_sync": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_i32,
                            c_i32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "pgd_peer_window_create": [c_vp, c_i64, c_vp],
    "pgd_peer_window_open": [c_vp, c_i32, c_i32, c_vp],
    "pgd_peer_window_destroy": [c_vp],
    "pgd_pcg_x0_sync": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_i32, c_i32, c_i32, c_i32, c_vp,
                        ctypes.POINTER(c_i32), ctypes.POINTER(c_dbl), c_vp],
    "pgd_pcg_persist_sync": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_dbl, c_dbl, c_i32, c_i32, c_vp, c_vp, c_vp,
                             c_vp, c_vp, c_vp, c_i32, ctypes.POINTER(c_i32), ctypes.POINTER(c_dbl), c_vp],
    "pgd_pcg_persist_start": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_dbl, c_dbl, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp],
    "pgd_get_phase_ns": [c_vp, c_vp, c_i32],
    "pgd_row_stats": [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp],
    "pgd_banded_solve": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp],
    "pgd_eval_weights": [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp],
    "pgd_eval_gemv": [c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp, c_vp],
    "pgd_eval_gemm_f64": [c_vp, c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i64, c_vp, c_i64, c_vp],
    "pgd_lincomb_dev": [c_vp, c_i32, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp],
    "pgd_scalar_programs": [c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp],
    "pgd_pcg_start": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_i32, c_i32, c_vp, c_i32, c_vp],
    "pgd_pcg_finish": [c_vp, ctypes.POINTER(c_i32), ctypes.POINTER(c_dbl)],
    "pgd_locate_points": [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i32, c_dbl, c_vp, c_vp, c_vp],
    "pgd_probe_modes": [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp],
}

_lib = None
_lock = threading.Lock()
_handles = {}
launch_count = 0  # kernels-launching C-ABI calls made through this module (for bench gpu_launches)


class PGDB200Error(RuntimeError):
    pass


def load_library():
    """dlopen libpgdb200.so; raises (never falls back) when it is missing."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise PGDB200Error(
                f"{LIB_PATH} is missing: build it with `python -m pgdrome_b200._build` "
                "(nvcc, sm_100a). pgdrome_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_char_p if name == "pgd_last_error" else c_i32
        _lib = lib
        return lib


_CUDA_OK = None


def require_cuda():
    global _CUDA_OK
    if _CUDA_OK is None:
        _CUDA_OK = bool(torch.cuda.is_available())
    if not _CUDA_OK:
        raise PGDB200Error("pgdrome_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def handle(device=None):
    """One library handle per device."""
    require_cuda()
    lib = _lib if _lib is not None else load_library()
    if device is None:
        dev = torch.cuda.current_device()
    elif isinstance(device, torch.device):
        dev = device.index if device.index is not None else torch.cuda.current_device()
    else:
        dev = torch.device(device).index
    if dev not in _handles:
        h = c_vp()
        rc = lib.pgd_create(dev, ctypes.byref(h))
        if rc != 0:
            raise PGDB200Error(f"pgd_create(device={dev}) failed with code {rc}")
        _handles[dev] = h
    return _handles[dev]


def _stream():
    """raw cudaStream_t of torch's current stream (the C-level getter: this is called for every launch)"""
    try:
        return c_vp(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:
        return c_vp(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=None, rows_strided=False):
    """device pointer of a contiguous CUDA tensor (or NULL).  rows_strided: a 2-D view whose rows are
    contiguous but spaced by a leading dimension (passed separately to the kernel) is accepted."""
    if t is None:
        return c_vp(0)
    if not t.is_cuda:
        raise PGDB200Error("expected a CUDA tensor (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise PGDB200Error(f"expected dtype {dtype}, got {t.dtype}")
    if rows_strided and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return c_vp(t.data_ptr())
    if not t.is_contiguous():
        raise PGDB200Error("expected a contiguous tensor")
    return c_vp(t.data_ptr())


def _check(rc, h, what):
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = load_library().pgd_last_error(h)
        raise PGDB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


F64, I32, I64 = torch.float64, torch.int32, torch.int64

# host <-> device traffic made by the package (bytes), for bench.py's e2e accounting
traffic = {"h2d": 0, "d2h": 0}


_STAGE_BYTES = 32 << 20  # page-locked staging block per direction (larger arrays are copied in pieces)
_stage = {}


def _staging(kind):
    """The package's page-locked host block for one direction ("h2d" / "d2h"), allocated on first use:
    (tensor, NumPy byte view of the same memory -- host-side copies are plain single-threaded memcpys)."""
    ent = _stage.get(kind)
    if ent is None:
        buf = torch.empty(_STAGE_BYTES, dtype=torch.uint8).pin_memory()
        ent = (buf, buf.numpy())
        _stage[kind] = ent
    return ent


def to_device(a, dtype=None):
    """NumPy array / tensor -> contiguous CUDA tensor (counted as host->device traffic).  The bytes travel through a
    page-locked staging block, so the copy is a real DMA from pinned memory instead of the driver's blocking
    pageable path."""
    require_cuda()
    if not isinstance(a, torch.Tensor):
        a = np.ascontiguousarray(a)
        if not a.flags.writeable:  # torch refuses to alias read-only arrays silently; the bytes are copied below anyway
            a = a.copy()
    t = torch.as_tensor(a, dtype=dtype).contiguous()
    nbytes = t.numel() * t.element_size()
    traffic["h2d"] += nbytes
    dev = torch.device("cuda", torch.cuda.current_device())
    if t.is_cuda or nbytes == 0 or t.is_pinned():
        return t.to(dev)
    out = torch.empty(t.shape, dtype=t.dtype, device=dev)
    src = t.view(-1).view(torch.uint8).numpy()
    dst = out.view(-1).view(torch.uint8)
    stage, stage_np = _staging("h2d")
    for o in range(0, nbytes, _STAGE_BYTES):
        m = min(_STAGE_BYTES, nbytes - o)
        stage_np[:m] = src[o:o + m]
        dst[o:o + m].copy_(stage[:m], non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the block is reused by the next piece / call
    return out


def to_host(t):
    """CUDA tensor -> NumPy array (counted as device->host traffic; synchronises).  Staged through page-locked memory."""
    nbytes = t.numel() * t.element_size()
    traffic["d2h"] += nbytes
    if not t.is_cuda or nbytes == 0 or nbytes > _STAGE_BYTES:
        return t.cpu().numpy()
    t = t.contiguous()
    stage, stage_np = _staging("d2h")
    stage[:nbytes].copy_(t.view(-1).view(torch.uint8), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return stage_np[:nbytes].copy().view(_NP_DTYPE[t.dtype]).reshape(tuple(t.shape))


_NP_DTYPE = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64,
             torch.uint8: np.uint8, torch.bool: np.bool_, torch.int8: np.int8, torch.int16: np.int16}


def stats(device=None, reset=False):
    """dict(launches, pcg_solves, pcg_iters, pcg_ms) from the library handle."""
    h, lib = handle(device), load_library()
    counts = (c_i64 * 4)()
    ms = c_dbl(0.0)
    lib.pgd_get_stats(h, ctypes.cast(counts, c_vp), ctypes.cast(ctypes.pointer(ms), c_vp), 1 if reset else 0)
    return {"launches": int(counts[0]), "pcg_solves": int(counts[1]), "pcg_iters": int(counts[2]),
            "pcg_resident_solves": int(counts[3]), "pcg_ms": float(ms.value)}


def set_option(name, value, device=None):
    h, lib = handle(device), load_library()
    rc = lib.pgd_set_option(h, name.encode(), int(value))
    if rc != 0:
        raise PGDB200Error("pgd_set_option(%s) failed: %s" % (name, lib.pgd_last_error(h).decode()))


# ------------------------------------------------------------------------------ multi-GPU
_COMM = {}


def comm_init(group=None):
    """Create the library's own NCCL communicator over the ranks of `group` (default: the world):
    rank 0 draws the unique id, torch.distributed carries it to the others.  Idempotent per device."""
    import torch.distributed as dist

    h, lib = handle(None), load_library()
    key = (torch.cuda.current_device(), id(group))
    if key in _COMM:
        return _COMM[key]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = (ctypes.c_ubyte * 128)()
    if rank == 0:
        rc = lib.pgd_comm_unique_id(ctypes.cast(buf, c_vp))
        if rc != 0:
            raise PGDB200Error("pgd_comm_unique_id failed (%d): NCCL not loadable" % rc)
    t = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    raw = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
    _check(lib.pgd_comm_init(h, ctypes.cast(raw, c_vp), rank, world), h, "pgd_comm_init")
    _COMM[key] = (rank, world)
    return _COMM[key]


_WINDOW = {}


def peer_window(n_local, group=None):
    """Collective: make sure this device has an NVLink peer window with room for p (n_local doubles on the
    largest rank) mapped on every rank of `group`; returns (rank, world) or None when P2P is unavailable."""
    import torch.distributed as dist

    h, lib = handle(None), load_library()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world > 16:
        return None
    dev = torch.cuda.current_device()
    need = torch.tensor([int(n_local)], dtype=I64, device="cuda")
    dist.all_reduce(need, op=dist.ReduceOp.MAX, group=group)
    need = int(need.item())
    key = (dev, id(group))
    if key in _WINDOW and _WINDOW[key] >= need:
        return rank, world
    cap = max(need + need // 8, 1024)
    if key in _WINDOW:
        dist.barrier(group=group)  # growing: no rank may free its old window while a peer can still be using it
    buf = (ctypes.c_ubyte * 64)()
    _check(lib.pgd_peer_window_create(h, cap, ctypes.cast(buf, c_vp)), h, "pgd_peer_window_create")
    mine = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    raw = (ctypes.c_ubyte * (64 * world))(*torch.cat(allh).cpu().tolist())
    try:
        _check(lib.pgd_peer_window_open(h, rank, world, ctypes.cast(raw, c_vp)), h, "pgd_peer_window_open")
        ok = 1
    except PGDB200Error:
        ok = 0
    flag = torch.tensor([ok], dtype=I64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:  # some rank could not map a peer: everybody falls back to NCCL
        lib.pgd_peer_window_destroy(h)
        _WINDOW.pop(key, None)
        return None
    _WINDOW[key] = cap
    return rank, world


def spcg_solve(A, b, rtol=1e-13, atol=0.0, maxit=10000, check_every=50, block=1, work=None):
    """Sharded PCG solve (pgd_spcg_solve_sync): A is a partition.ShardedMatrix, b the owned slice.
    Needs comm_init() when more than one rank takes part."""
    h, lib = handle(b.device), load_library()
    halo = A.halo
    world = len(halo.send_counts)
    n_send = int(sum(halo.send_counts))
    need = A.n_owned * (3 + block) + A.n_local + n_send + 8
    if work is None or work.numel() < need:
        work = torch.empty(need, dtype=F64, device=b.device)
    x = torch.empty(A.n_owned, dtype=F64, device=b.device)
    sc = (c_i64 * world)(*[int(v) for v in halo.send_counts])
    rc_ = (c_i64 * world)(*[int(v) for v in halo.recv_counts])
    iters, relres = c_i32(0), c_dbl(0.0)
    gb = getattr(halo, "peer_ghost_base", None)
    gbase = (c_i64 * world)(*[int(v) for v in gb]) if gb is not None else None
    _check(lib.pgd_spcg_solve_sync(h, _p(A.rowptr, I32), _p(A.colidx, I32), _p(A.values, F64), _p(b, F64), _p(x), A.n_owned,
                                   A.n_local, block, _p(halo.send_idx, I64) if n_send else c_vp(0), ctypes.cast(sc, c_vp),
                                   ctypes.cast(rc_, c_vp), float(rtol), float(atol), int(maxit), int(check_every), _p(work),
                                   ctypes.cast(ctypes.pointer(iters), c_vp), ctypes.cast(ctypes.pointer(relres), c_vp),
                                   _stream(), ctypes.cast(gbase, c_vp) if gbase is not None else c_vp(0)), h,
           "pgd_spcg_solve_sync")
    return x, int(iters.value), float(relres.value)


# ------------------------------------------------------------------------------ pattern
def pattern_build(cell_dofs, n_dofs):
    """cell_dofs int32 [n_cells, ndl] (cuda) -> rowptr, colidx, gptr, gidx."""
    h, lib = handle(cell_dofs.device), load_library()
    n_cells, ndl = cell_dofs.shape
    nnz = c_i64(0)
    _check(lib.pgd_pattern_build_sync(h, _p(cell_dofs, I32), n_cells, ndl, n_dofs, ctypes.byref(nnz), _stream()), h,
           "pgd_pattern_build_sync")
    dev = cell_dofs.device
    rowptr = torch.empty(n_dofs + 1, dtype=I32, device=dev)
    colidx = torch.empty(nnz.value, dtype=I32, device=dev)
    gptr = torch.empty(nnz.value + 1, dtype=I64, device=dev)
    gidx = torch.empty(n_cells * ndl * ndl, dtype=I32, device=dev)
    _check(lib.pgd_pattern_export(h, _p(rowptr), _p(colidx), _p(gptr), _p(gidx), _stream()), h, "pgd_pattern_export")
    return rowptr, colidx, gptr, gidx


def vecmap_build(cell_dofs, n_dofs):
    h, lib = handle(cell_dofs.device), load_library()
    n_cells, ndl = cell_dofs.shape
    vptr = torch.empty(n_dofs + 1, dtype=I64, device=cell_dofs.device)
    vidx = torch.empty(n_cells * ndl, dtype=I32, device=cell_dofs.device)
    _check(lib.pgd_vecmap_build_sync(h, _p(cell_dofs, I32), n_cells, ndl, n_dofs, _p(vptr), _p(vidx), _stream()), h,
           "pgd_vecmap_build_sync")
    return vptr, vidx


# ------------------------------------------------------------------------------ assembly
def elem_bilinear(coords, cell_verts, tdim, gdim, bs, nd, phi, dphi, qw, wq, T, out=None):
    h, lib = handle(coords.device), load_library()
    n_cells = cell_verts.shape[0]
    nq = qw.numel()
    ndl = nd * bs
    if out is None:
        out = torch.empty(n_cells * ndl * ndl, dtype=F64, device=coords.device)
    _check(lib.pgd_elem_bilinear(h, _p(coords, F64), _p(cell_verts, I32), n_cells, tdim, gdim, bs, nd, nq, _p(phi, F64),
                                 _p(dphi, F64), _p(qw, F64), _p(wq, F64) if wq is not None else c_vp(0), _p(T, F64),
                                 _p(out), _stream()), h, "pgd_elem_bilinear")
    return out


def elem_linear(coords, cell_verts, tdim, gdim, bs, nd, phi, dphi, qw, wq, L, out=None):
    h, lib = handle(coords.device), load_library()
    n_cells = cell_verts.shape[0]
    nq = qw.numel()
    if out is None:
        out = torch.empty(n_cells * nd * bs, dtype=F64, device=coords.device)
    _check(lib.pgd_elem_linear(h, _p(coords, F64), _p(cell_verts, I32), n_cells, tdim, gdim, bs, nd, nq, _p(phi, F64),
                               _p(dphi, F64), _p(qw, F64), _p(wq, F64) if wq is not None else c_vp(0), _p(L, F64),
                               _p(out), _stream()), h, "pgd_elem_linear")
    return out


def gather_values(src, gptr, gidx, n_out, out=None):
    h, lib = handle(src.device), load_library()
    if out is None:
        out = torch.empty(n_out, dtype=F64, device=src.device)
    _check(lib.pgd_gather_values(h, _p(src, F64), _p(gptr, I64), _p(gidx, I32), n_out, _p(out), _stream()), h,
           "pgd_gather_values")
    return out


def assemble_p1(coords, cell_verts, gdim, c_mass, c_stiff, c_adv, gptr, gidx, nnz, out=None):
    h, lib = handle(coords.device), load_library()
    if out is None:
        out = torch.empty(nnz, dtype=F64, device=coords.device)
    adv = None
    if c_adv is not None:
        adv = (c_dbl * 3)(*[float(v) for v in list(c_adv) + [0.0] * (3 - len(c_adv))])
    _check(lib.pgd_assemble_p1(h, _p(coords, F64), _p(cell_verts, I32), cell_verts.shape[0], gdim, float(c_mass),
                               float(c_stiff), ctypes.cast(adv, c_vp) if adv is not None else c_vp(0), _p(gptr, I64),
                               _p(gidx, I32), nnz, _p(out), _stream()), h, "pgd_assemble_p1")
    return out


def p1_rowplan_build(rowptr, colidx, cell_dofs, vptr, vidx, n_nodes):
    """Plan of the row-owner fused P1 kernel: int32 [n_cells * nv * 2] (pairs {vidx entry, packed positions})."""
    h, lib = handle(rowptr.device), load_library()
    n_cells, nv = cell_dofs.shape
    vent = torch.empty(n_cells * nv * 2, dtype=I32, device=rowptr.device)
    _check(lib.pgd_p1_rowplan_build_sync(h, _p(rowptr, I32), _p(colidx, I32), _p(cell_dofs, I32), n_cells, nv, _p(vptr, I64),
                                         _p(vidx, I32), n_nodes, _p(vent), _stream()), h, "pgd_p1_rowplan_build_sync")
    return vent


def assemble_p1_rows(coords, cell_verts, gdim, c_mass, c_stiff, c_adv, rowptr, vptr, vent, n_nodes, out=None, coords_soa=None,
                     nnz=None):
    """coords_soa: optional component-major copy [gdim, n_verts] of coords (faster coordinate gathers).  nnz: size of the
    value array (known to the owner of the pattern; read back from rowptr[-1] when neither it nor ``out`` is given)."""
    h, lib = handle(coords.device), load_library()
    if out is None:
        out = torch.empty(int(rowptr[-1].item()) if nnz is None else int(nnz), dtype=F64, device=coords.device)
    adv = None
    if c_adv is not None:
        adv = (c_dbl * 3)(*[float(v) for v in list(c_adv) + [0.0] * (3 - len(c_adv))])
    _check(lib.pgd_assemble_p1_rows(h, _p(coords, F64), _p(cell_verts, I32), cell_verts.shape[0], gdim, float(c_mass),
                                    float(c_stiff), ctypes.cast(adv, c_vp) if adv is not None else c_vp(0), _p(rowptr, I32),
                                    _p(vptr, I64), _p(vent, I32), n_nodes, _p(out),
                                    _p(coords_soa, F64) if coords_soa is not None else c_vp(0),
                                    coords_soa.shape[1] if coords_soa is not None else 0, _stream()), h, "pgd_assemble_p1_rows")
    return out


def assemble_p1_tensor(gdim, bs, T, coords_soa, cell_verts, rowptr, node_vptr, node_vent, n_rows, nnz, w_cell=None, out=None):
    """Fused row-owner assembly of a constant-coefficient P1 atom with form tensor T (pgd_assemble_p1_tensor)."""
    h, lib = handle(rowptr.device), load_library()
    if out is None:
        out = torch.empty(int(nnz), dtype=F64, device=rowptr.device)
    Th = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(bs, gdim + 1, bs, gdim + 1))
    _check(lib.pgd_assemble_p1_tensor(h, gdim, bs, Th.ctypes.data_as(c_vp), _p(coords_soa, F64), coords_soa.shape[1],
                                      _p(cell_verts, I32), _p(rowptr, I32), _p(node_vptr, I64), _p(node_vent, I32),
                                      _p(w_cell, F64) if w_cell is not None else c_vp(0), int(n_rows), _p(out, F64), _stream()),
           h, "pgd_assemble_p1_tensor")
    return out


def lincomb(xs, coefs, out=None, accumulate=False):
    """out = (out if accumulate) + sum_t coefs[t] * xs[t]; xs: list of equally sized CUDA tensors."""
    n_terms = len(xs)
    ref = out if out is not None else xs[0]
    h, lib = handle(ref.device), load_library()
    n = ref.numel()
    if out is None:
        out = torch.empty(n, dtype=F64, device=ref.device)
    ptrs = (c_vp * max(1, n_terms))(*[_p(x, F64).value for x in xs])
    if isinstance(coefs, torch.Tensor):  # coefficients computed on the device (scalar_programs): no host round trip
        _check(lib.pgd_lincomb_dev(h, n_terms, ctypes.cast(ptrs, c_vp), _p(coefs, F64), n, _p(out, F64),
                                   1 if accumulate else 0, _stream()), h, "pgd_lincomb_dev")
        return out
    cs = (c_dbl * max(1, n_terms))(*[float(c) for c in coefs])
    _check(lib.pgd_lincomb(h, n_terms, ctypes.cast(ptrs, c_vp), ctypes.cast(cs, c_vp), n, _p(out, F64),
                           1 if accumulate else 0, _stream()), h, "pgd_lincomb")
    return out


SP_MAX_PROG, SP_MAX_CODE, SP_MAX_CONST = 32, 640, 128
OP_CONST, OP_LOAD, OP_MUL, OP_ADD, OP_SUB, OP_DIV, OP_NEG = range(7)


def scalar_programs(programs, consts, pool, out):
    """programs: list of postfix instruction lists [(opcode << 24) | operand, ...] (pgd_scalar_programs); consts: list of
    floats; pool: device tensor the LOAD operands index; out: device tensor [len(programs)] (filled in place, chunked
    to the 32-program / 640-instruction limit of one call)."""
    h, lib = handle(out.device), load_library()
    g0 = 0
    cs = (c_dbl * max(1, len(consts)))(*consts)
    while g0 < len(programs):
        g1, n_code = g0, 0
        while g1 < len(programs) and g1 - g0 < SP_MAX_PROG and n_code + len(programs[g1]) <= SP_MAX_CODE:
            n_code += len(programs[g1])
            g1 += 1
        if g1 == g0:
            raise ValueError("scalar program longer than %d instructions" % SP_MAX_CODE)
        off, code = [0], []
        for p in programs[g0:g1]:
            code.extend(p)
            off.append(len(code))
        a_off = (c_i32 * len(off))(*off)
        a_code = (c_i32 * max(1, len(code)))(*code)
        _check(lib.pgd_scalar_programs(h, g1 - g0, ctypes.cast(a_off, c_vp), ctypes.cast(a_code, c_vp), ctypes.cast(cs, c_vp),
                                       len(consts), _p(pool, F64) if pool is not None else c_vp(0),
                                       c_vp(out.data_ptr() + 8 * g0), _stream()), h, "pgd_scalar_programs")
        g0 = g1
    return out


# ------------------------------------------------------------------------------ BC / sparse
def apply_dirichlet(rowptr, colidx, values, b, bc_dofs, bc_vals=None):
    if bc_dofs is None or bc_dofs.numel() == 0:
        return
    h, lib = handle(rowptr.device), load_library()
    n = rowptr.numel() - 1
    work = torch.empty(2 * n, dtype=F64, device=rowptr.device) if (bc_vals is not None and values is not None and b is not None) else None
    _check(lib.pgd_apply_dirichlet(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64) if values is not None else c_vp(0),
                                   _p(b, F64) if b is not None else c_vp(0), _p(bc_dofs, I32),
                                   _p(bc_vals, F64) if bc_vals is not None else c_vp(0), bc_dofs.numel(), n,
                                   _p(work, F64) if work is not None else c_vp(0), _stream()), h,
           "pgd_apply_dirichlet")


def set_entries(x, idx, vals=None):
    if idx is None or idx.numel() == 0:
        return
    h, lib = handle(x.device), load_library()
    _check(lib.pgd_set_entries(h, _p(x, F64), _p(idx, I32), _p(vals, F64) if vals is not None else c_vp(0), idx.numel(),
                               _stream()), h, "pgd_set_entries")


def spmv(rowptr, colidx, values, x, y=None, lpr=0):
    h, lib = handle(x.device), load_library()
    n = rowptr.numel() - 1
    if y is None:
        y = torch.empty(n, dtype=F64, device=x.device)
    _check(lib.pgd_spmv(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(x, F64), _p(y, F64), n, lpr, _stream()), h,
           "pgd_spmv")
    return y


def spmv_dot(rowptr, colidx, values, x, w, y=None, out=None, lpr=0):
    h, lib = handle(x.device), load_library()
    n = rowptr.numel() - 1
    if y is None:
        y = torch.empty(n, dtype=F64, device=x.device)
    if out is None:
        out = torch.empty(1, dtype=F64, device=x.device)
    _check(lib.pgd_spmv_dot(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(x, F64), _p(y, F64), _p(w, F64),
                            _p(out, F64), n, lpr, _stream()), h, "pgd_spmv_dot")
    return y, out


def bilinear(rowptr, colidx, values, x, y, out=None, lpr=0):
    """out[0] = x^T A y (device scalar)."""
    h, lib = handle(x.device), load_library()
    n = rowptr.numel() - 1
    if out is None:
        out = torch.empty(1, dtype=F64, device=x.device)
    _check(lib.pgd_bilinear(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(x, F64), _p(y, F64), n, _p(out, F64),
                            lpr, _stream()), h, "pgd_bilinear")
    return out


def dot(x, y, out=None):
    h, lib = handle(x.device), load_library()
    if out is None:
        out = torch.empty(1, dtype=F64, device=x.device)
    _check(lib.pgd_dot(h, _p(x, F64), _p(y, F64), x.numel(), _p(out, F64), _stream()), h, "pgd_dot")
    return out


def panel_dots(P, n_vecs, x, out=None):
    """P [>=n_vecs, ld] row-major; out[m] = P[m,:n] . x."""
    h, lib = handle(x.device), load_library()
    if out is None:
        out = torch.empty(n_vecs, dtype=F64, device=x.device)
    _check(lib.pgd_panel_dots(h, _p(P, F64), P.stride(0) if P.dim() == 2 else x.numel(), n_vecs, _p(x, F64), x.numel(),
                              _p(out, F64), _stream()), h, "pgd_panel_dots")
    return out


# ------------------------------------------------------------------------------ solves
def pcg(rowptr, colidx, values, b, x=None, rtol=1e-12, atol=0.0, maxit=20000, check_every=50, block=1, lpr=0, work=None,
        x0=None):
    """Jacobi / node-block-Jacobi PCG.  x0: optional initial guess (copied; warm start, pgd_pcg_x0_sync)."""
    h, lib = handle(b.device), load_library()
    n = b.numel()
    if x0 is not None:
        x = x0.detach().clone() if x is None else x.copy_(x0)
    elif x is None:
        x = torch.empty(n, dtype=F64, device=b.device)
    if work is None or work.numel() < (5 + block) * n + 8:
        work = torch.empty((5 + block) * n + 8, dtype=F64, device=b.device)
    iters, relres = c_i32(0), c_dbl(0.0)
    fn = lib.pgd_pcg_x0_sync if x0 is not None else lib.pgd_pcg_sync
    _check(fn(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(b, F64), _p(x, F64), n, rtol, atol,
              maxit, check_every, block, lpr, _p(work, F64), ctypes.byref(iters), ctypes.byref(relres),
              _stream()), h, "pgd_pcg_sync")
    return x, iters.value, relres.value


def bsr_plan(rowptr, colidx, bs):
    """Block-column list of a node-blocked CSR pattern for the node-block walk of pgd_pcg_persist_sync: (bcol int32
    [nnz / bs^2], longest block row), or None when the pattern is not a union of full bs x bs blocks."""
    n = rowptr.numel() - 1
    if bs < 2 or n % bs:
        return None
    rp = rowptr.to(I64)
    lens = rp[1:] - rp[:-1]
    if bool((lens.view(-1, bs) != lens.view(-1, bs)[:, :1]).any()) or bool((lens % bs != 0).any()):
        return None
    # entries of the first row of every node whose column is a first component: one per block, in block order
    first = torch.repeat_interleave(torch.arange(n, device=rowptr.device) % bs == 0, lens)
    sel = first & (colidx % bs == 0)
    bcol = torch.div(colidx[sel], bs, rounding_mode="floor").to(I32).contiguous()
    nnz = int(colidx.numel())
    if bcol.numel() * bs * bs != nnz:
        return None
    # every block must be full: columns of a row = bs * bcol + (0..bs-1), identical for the bs rows of a node
    exp = (bcol.to(I64).repeat_interleave(bs) * bs + torch.arange(bs, device=bcol.device).repeat(bcol.numel()))
    nb = (lens.view(-1, bs)[:, 0] // bs)
    for i in range(bs):
        rows_i = torch.arange(i, n, bs, device=rowptr.device)
        seg = torch.repeat_interleave(rp[rows_i], lens[rows_i]) + _ramp(lens[rows_i])
        if not torch.equal(colidx[seg].to(I64), exp):
            return None
    return bcol, int(nb.max().item())


def _ramp(lens):
    """0..len-1 for every segment, concatenated (device)."""
    ends = torch.cumsum(lens, 0)
    return torch.arange(int(ends[-1].item()), device=lens.device) - torch.repeat_interleave(ends - lens, lens)


def pcg_persist(rowptr, colidx, values, b, x=None, n_owned=None, block=1, rtol=1e-13, atol=0.0, maxit=20000, x0=None,
                halo=None, bsr=None, work=None, defer=False):
    """Persistent streaming PCG (pgd_pcg_persist_sync).  rowptr: the n_owned local rows; x / x0: [n_local] (ghosts of
    x0 valid); halo: partition.HaloPlan with its peer layout (sharded solves) or None; bsr: result of ``bsr_plan``.
    Returns (x [n_local], iterations, relative residual).  defer=True (single GPU only): pgd_pcg_persist_start -- the
    solve is only enqueued, x is valid in stream order, iterations = -1; collect with ``pcg_finish``."""
    h, lib = handle(b.device), load_library()
    no = rowptr.numel() - 1 if n_owned is None else int(n_owned)
    nl = no if halo is None else int(halo.n_local)
    if x is None:
        x = torch.empty(nl, dtype=F64, device=b.device)
    if x0 is not None and x0.data_ptr() != x.data_ptr():
        x.copy_(x0)
    need = 3 * ((no + 1) & ~1) + ((no * block + 1) & ~1) + 2 * ((nl + 1) & ~1) + 8
    if work is None or work.numel() < need:
        work = torch.empty(need, dtype=F64, device=b.device)
    send_idx = sc = rc_ = gbase = None
    if halo is not None:
        world = len(halo.send_counts)
        sc = (c_i64 * world)(*[int(v) for v in halo.send_counts])
        rc_ = (c_i64 * world)(*[int(v) for v in halo.recv_counts])
        gbase = (c_i64 * world)(*[int(v) for v in halo.peer_ghost_base])
        send_idx = halo.send_idx if halo.send_idx.numel() else None
    iters, relres = c_i32(0), c_dbl(0.0)
    cast = lambda a: ctypes.cast(a, c_vp) if a is not None else c_vp(0)
    if defer:
        if halo is not None:
            raise ValueError("deferred persistent solves are single-GPU only")
        _check(lib.pgd_pcg_persist_start(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(b, F64), _p(x, F64), no, block,
                                         float(rtol), float(atol), int(maxit), 1 if x0 is not None else 0, _p(work, F64),
                                         _p(bsr[0], I32) if bsr else c_vp(0), int(bsr[1]) if bsr else 0, _stream()), h,
               "pgd_pcg_persist_start")
        return x, -1, 0.0
    rc = lib.pgd_pcg_persist_sync(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(b, F64), _p(x, F64), no, nl, block,
                                  float(rtol), float(atol), int(maxit), 1 if x0 is not None else 0, _p(work, F64),
                                  _p(send_idx, I64), cast(sc), cast(rc_), cast(gbase),
                                  _p(bsr[0], I32) if bsr else c_vp(0), int(bsr[1]) if bsr else 0,
                                  ctypes.byref(iters), ctypes.byref(relres), _stream())
    if rc == -6:
        # a peer did not arrive: the library has disabled its window (sequence numbers may have diverged); forget the
        # cached one so that the next sharded solve creates and opens a fresh window collectively
        _WINDOW.clear()
    _check(rc, h, "pgd_pcg_persist_sync")
    return x, iters.value, relres.value


def phase_ns(reset=True, device=None):
    """per-phase nanoseconds of the persistent PCG kernel (set_option("prof", 1)): D, barrier, S, reduce1, U, reduce2"""
    h, lib = handle(device), load_library()
    ns = (c_i64 * 6)()
    _check(lib.pgd_get_phase_ns(h, ctypes.cast(ns, c_vp), 1 if reset else 0), h, "pgd_get_phase_ns")
    return dict(zip(["direction", "barrier", "spmv", "reduce_pq", "update", "reduce_rz"], [int(v) for v in ns]))


def pcg_start(rowptr, colidx, values, b, rtol=1e-12, atol=0.0, maxit=20000, block=1, x0=None):
    """Enqueue the SM-resident PCG without waiting (pgd_pcg_start).  Returns the solution tensor (valid in stream
    order) or None when the system is not known to fit the resident solver -- then call ``pcg``.  Collect iterations
    and residual with ``pcg_finish`` before the next solve."""
    h, lib = handle(b.device), load_library()
    n = b.numel()
    x = x0.detach().clone() if x0 is not None else torch.empty(n, dtype=F64, device=b.device)
    work = torch.empty((5 + block) * n + 8, dtype=F64, device=b.device)
    rc = lib.pgd_pcg_start(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(b, F64), _p(x, F64), n, rtol, atol,
                           maxit, block, _p(work, F64), 1 if x0 is not None else 0, _stream())
    if rc == 1:
        return None
    _check(rc, h, "pgd_pcg_start")
    return x


def pcg_finish(device=None):
    """(iterations, relative residual) of the solve started with ``pcg_start``; (-1, 0.0) if none is pending."""
    h, lib = handle(device), load_library()
    iters, relres = c_i32(0), c_dbl(0.0)
    _check(lib.pgd_pcg_finish(h, ctypes.byref(iters), ctypes.byref(relres)), h, "pgd_pcg_finish")
    return iters.value, relres.value


def banded_solve(rowptr, colidx, values, b, perm, kl, ku, x=None, work=None, info=None):
    h, lib = handle(b.device), load_library()
    n = b.numel()
    if x is None:
        x = torch.empty(n, dtype=F64, device=b.device)
    need = (2 * kl + ku + 3) * n
    if work is None or work.numel() < need:
        work = torch.empty(need, dtype=F64, device=b.device)
    if info is None:
        info = torch.zeros(1, dtype=I32, device=b.device)
    _check(lib.pgd_banded_solve(h, _p(rowptr, I32), _p(colidx, I32), _p(values, F64), _p(b, F64), _p(x, F64), n,
                                _p(perm, I32), kl, ku, _p(work, F64), _p(info, I32), _stream()), h, "pgd_banded_solve")
    return x, info


# ------------------------------------------------------------------------------ evaluate
def eval_weights(xs, cds, Phis, degs, R, points, out=None):
    """xs[i] [nc_i+1] f64, cds[i] [nc_i, deg_i+1] i32, Phis[i] [>=R, ld_i] f64, points [C, n_free]."""
    n_free = len(xs)
    dev = points.device
    h, lib = handle(dev), load_library()
    C = points.shape[0]
    if out is None:
        out = torch.empty((R, C), dtype=F64, device=dev)
    flag = torch.zeros(1, dtype=I32, device=dev)
    m = max(1, n_free)
    a_xs = (c_vp * m)(*[_p(t, F64).value for t in xs])
    a_cd = (c_vp * m)(*[_p(t, I32).value for t in cds])
    a_ph = (c_vp * m)(*[_p(t, F64).value for t in Phis])
    a_nc = (c_i32 * m)(*[t.numel() - 1 for t in xs])
    a_dg = (c_i32 * m)(*[int(d) for d in degs])
    a_ld = (c_i64 * m)(*[t.stride(0) for t in Phis])
    _check(lib.pgd_eval_weights(h, n_free, ctypes.cast(a_xs, c_vp), ctypes.cast(a_cd, c_vp), ctypes.cast(a_ph, c_vp),
                                ctypes.cast(a_nc, c_vp), ctypes.cast(a_dg, c_vp), ctypes.cast(a_ld, c_vp), R,
                                _p(points, F64), C, _p(out, F64), _p(flag, I32), _stream()), h, "pgd_eval_weights")
    return out, flag


def eval_gemv(X, R, w, out=None):
    """X [>=R, ld] row-major modes; out[n] = sum_k X[k, n] w[k]."""
    h, lib = handle(X.device), load_library()
    N = X.shape[1]
    if out is None:
        out = torch.empty(N, dtype=F64, device=X.device)
    _check(lib.pgd_eval_gemv(h, _p(X, F64), X.stride(0), R, _p(w, F64), N, _p(out, F64), _stream()), h, "pgd_eval_gemv")
    return out


def locate_points(coords, cells, points, tol=1e-10):
    """coords [n_verts, g] f64, cells [n_cells, g+1] i32, points [n_pts, g] f64 (device) -> (cell i32 [n_pts]
    with -1 = outside, bary f64 [n_pts, g+1]); the lowest-numbered containing cell wins."""
    h, lib = handle(coords.device), load_library()
    g = coords.shape[1]
    n_pts = points.shape[0]
    cell = torch.empty(n_pts, dtype=I32, device=coords.device)
    bary = torch.empty((n_pts, g + 1), dtype=F64, device=coords.device)
    _check(lib.pgd_locate_points(h, _p(coords, F64), _p(cells, I32), cells.shape[0], g, _p(points, F64), n_pts, float(tol),
                                 _p(cell, I32), _p(bary, F64), _stream()), h, "pgd_locate_points")
    return cell, bary


def probe_modes(X, R, dofs, w):
    """X [>=R, ld] modes, dofs i32 [n_rows, nd], w f64 [n_rows, nd] -> E [R, n_rows], E[k, r] = sum_j w[r, j] X[k, dofs[r, j]]."""
    h, lib = handle(X.device), load_library()
    n_rows, nd = dofs.shape
    E = torch.empty((R, n_rows), dtype=F64, device=X.device)
    _check(lib.pgd_probe_modes(h, _p(X, F64, True), X.stride(0), R, _p(dofs, I32), _p(w, F64), n_rows, nd, _p(E, F64), n_rows,
                               _stream()), h, "pgd_probe_modes")
    return E


ROW_STATS = ("min", "max", "min_abs", "max_abs", "sum_sq", "sum_sq_diff", "sum_sq_ref")


def row_stats(U, F=None):
    """U [C, N] (rows contiguous), optional reference rows F [C, N] -> device tensor [C, 7] (see ROW_STATS)."""
    h, lib = handle(U.device), load_library()
    C, N = U.shape
    out = torch.empty((C, len(ROW_STATS)), dtype=F64, device=U.device)
    _check(lib.pgd_row_stats(h, _p(U, F64, True), U.stride(0), _p(F, F64, True) if F is not None else c_vp(0),
                             F.stride(0) if F is not None else 0, C, N, _p(out, F64), _stream()), h, "pgd_row_stats")
    return out


def eval_gemm(W, X, R, out=None):
    """W [>=R, C], X [>=R, N] -> U [C, N] = W^T X on the FP64 tensor cores."""
    h, lib = handle(X.device), load_library()
    C, N = W.shape[1], X.shape[1]
    if out is None:
        out = torch.empty((C, N), dtype=F64, device=X.device)
    _check(lib.pgd_eval_gemm_f64(h, _p(W, F64, True), W.stride(0), _p(X, F64, True), X.stride(0), R, C, N, _p(out, F64, True),
                                 out.stride(0),
                                 _stream()), h, "pgd_eval_gemm_f64")
    return out
