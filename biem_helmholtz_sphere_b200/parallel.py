"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

The path shards three natural ways and none of them has an exchange step:

* wavenumber sweeps   -- system ``i`` goes to rank ``i mod world``; nothing is communicated while solving, the per-k
  outputs (density 64 KB per k at C3, probe values) are all-gathered afterwards only if the caller wants them everywhere;
* field evaluation    -- contiguous row tiles of the point grid per rank, after ONE broadcast of the solved density
  (590 KB at C5) from the rank that solved the system: the only collective on the data path;
* assembly of ONE large system -- block rows (all harmonics of a row ball) per rank, then an all-gather of the strips
  onto every rank (21.7 GB at C5, ~25 ms over NVLink): optional, the single-GPU assembly is already 0.2 % of the solve;
* the LU of one system does not shard (replicas only).

Everything here is host logic over tensors that live wherever the process group's backend wants them (CUDA for NCCL, CPU
for gloo), so that it can be exercised with ``gloo`` and world_size 2 on a CPU-only box (tests/test_parallel_cpu.py).
The numerical work is injected through ``solve_shard`` / ``eval_tile`` callables; the defaults call the CUDA path.
"""

from __future__ import annotations

from collections.abc import Callable
from typing import Any

import numpy as np
import torch
import torch.distributed as dist


def dist_info(group=None) -> tuple[int, int]:
    """(rank, world) of the default / given process group; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(K: int, rank: int, world: int) -> np.ndarray:
    """Indices of the sweep owned by ``rank``: i = rank, rank + world, ... (round-robin keeps the shards balanced in
    cost when the work per system varies smoothly with k)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return np.arange(rank, K, world)


def field_rows(n_rows: int, rank: int, world: int) -> slice:
    """Contiguous block of grid rows evaluated by ``rank`` (sizes differ by at most one row)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def _comm_device(group=None) -> torch.device:
    backend = dist.get_backend(group)
    if "nccl" in str(backend):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _as_real(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


def all_gather_round_robin(local: torch.Tensor, K: int, group=None) -> torch.Tensor:
    """Inverse of :func:`shard_indices`: every rank passes its ``[K_local, ...]`` shard, every rank gets ``[K, ...]``.

    Shards may differ in length by one; they are padded to the longest for the collective."""
    rank, world = dist_info(group)
    if world == 1:
        return local
    kmax = -(-K // world)
    dev = _comm_device(group)
    buf = torch.zeros((kmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    buf[: local.shape[0]] = local.to(dev)
    parts = [torch.empty_like(_as_real(buf)) for _ in range(world)]
    dist.all_gather(parts, _as_real(buf).contiguous(), group=group)
    out = torch.empty((K,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    for r in range(world):
        idx = shard_indices(K, r, world)
        p = parts[r]
        p = torch.view_as_complex(p) if local.is_complex() else p
        out[torch.as_tensor(idx, device=dev)] = p[: len(idx)]
    return out.to(local.device)


def all_gather_rows(local: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """Inverse of :func:`field_rows` along dim 0."""
    rank, world = dist_info(group)
    if world == 1:
        return local
    rmax = -(-n_rows // world)
    dev = _comm_device(group)
    buf = torch.zeros((rmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    buf[: local.shape[0]] = local.to(dev)
    parts = [torch.empty_like(_as_real(buf)) for _ in range(world)]
    dist.all_gather(parts, _as_real(buf).contiguous(), group=group)
    out = []
    for r in range(world):
        sl = field_rows(n_rows, r, world)
        p = parts[r]
        p = torch.view_as_complex(p) if local.is_complex() else p
        out.append(p[: sl.stop - sl.start])
    return torch.cat(out, dim=0).to(local.device)


def broadcast_density(density: torch.Tensor | None, shape: tuple[int, ...], src: int = 0, group=None,
                      device: torch.device | None = None) -> torch.Tensor:
    """Broadcast the solved coefficients ``[B, H]`` (complex128) from ``src`` to every rank -- the one collective of
    the field-evaluation path.  Non-source ranks pass ``None``."""
    rank, world = dist_info(group)
    if world == 1:
        assert density is not None
        return density
    dev = _comm_device(group)
    if rank == src:
        assert density is not None and tuple(density.shape) == tuple(shape)
        buf = density.to(dev).to(torch.complex128).contiguous()
    else:
        buf = torch.empty(shape, dtype=torch.complex128, device=dev)
    r = torch.view_as_real(buf)
    dist.broadcast(r, src=src, group=group)
    return buf if device is None else buf.to(device)


# ---- default numerical back ends (CUDA path) ----------------------------------------------------------------
def _default_solve_shard(c, centers, radii, ks, n_end, eta, direction, x):
    from . import _biem

    dev = torch.device("cuda", torch.cuda.current_device())
    kk = torch.as_tensor(np.asarray(ks), dtype=torch.float64, device=dev)
    cen = torch.as_tensor(np.asarray(centers), dtype=torch.float64, device=dev)
    rad = torch.as_tensor(np.asarray(radii), dtype=torch.float64, device=dev)
    d = cen.shape[-1]
    dr = torch.as_tensor(np.asarray(direction), dtype=torch.float64, device=dev).reshape(d, 1)
    et = torch.full_like(kk, float(eta))
    uin, _ = _biem.plane_wave(k=kk, direction=dr)
    res = _biem.biem(c, centers=cen[None], radii=rad[None], k=kk, n_end=n_end, eta=et, uin=uin, keep_matrix=False)
    u = None
    if x is not None:
        u = res.uscat(torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)).movedim(-1, 0)  # [K_local, P]
    return res.density, u


def _default_eval_tile(c, centers, radii, k, eta, n_end, density, x_tile):
    from . import _ops
    from ._coords import tree_spec

    dev = torch.device("cuda", torch.cuda.current_device())
    d = x_tile.shape[0]
    axes = list(tree_spec(c).axes)  # the device code works in the chain frame of the tree
    xt = torch.as_tensor(np.asarray(x_tile)[axes], dtype=torch.float64, device=dev).reshape(d, -1).contiguous()
    return _ops.uscat(d, n_end, torch.as_tensor(np.asarray(centers)[:, axes], device=dev),
                      torch.as_tensor(np.asarray(radii), device=dev), complex(k), float(eta), density.to(dev), xt)


# ---- sharded drivers --------------------------------------------------------------------------------------------
def sweep(c: Any, *, centers, radii, ks, n_end: int, eta: float = 1.0, direction=None, x=None, gather: bool = True,
          group=None, solve_shard: Callable | None = None) -> dict:
    """Wavenumber sweep over one geometry, sharded ``i -> rank i mod world`` with no communication while solving.

    Returns ``{"indices", "density" [K*, B, H], "uscat" [K*, P] | None}``; ``K*`` is the whole sweep when ``gather`` (an
    all-gather of the per-k outputs at the end) and this rank's shard otherwise."""
    rank, world = dist_info(group)
    ks = np.asarray(ks, dtype=np.float64)
    K = ks.shape[0]
    idx = shard_indices(K, rank, world)
    d = np.asarray(centers).shape[-1]
    if direction is None:
        direction = np.eye(d)[0]
    fn = solve_shard or _default_solve_shard
    if len(idx):
        dens, u = fn(c, centers, radii, ks[idx], n_end, eta, direction, x)
    else:  # more ranks than systems
        dens, u = None, None
    if not gather or world == 1:
        return {"indices": idx, "density": dens, "uscat": u}
    # shapes of an empty shard are not known locally: agree on them through the group
    meta = [None] * world
    dist.all_gather_object(meta, None if dens is None else (tuple(dens.shape[1:]), None if u is None else tuple(u.shape[1:])),
                           group=group)
    shp = next(m for m in meta if m is not None)
    dev = _comm_device(group)
    if dens is None:
        dens = torch.zeros((0,) + shp[0], dtype=torch.complex128, device=dev)
        u = None if shp[1] is None else torch.zeros((0,) + shp[1], dtype=torch.complex128, device=dev)
    out = {"indices": np.arange(K), "density": all_gather_round_robin(torch.as_tensor(dens), K, group), "uscat": None}
    if u is not None:
        out["uscat"] = all_gather_round_robin(torch.as_tensor(u), K, group)
    return out


def uscat_sharded(c: Any, *, centers, radii, k: float, eta: float, n_end: int, density: torch.Tensor | None,
                  density_shape: tuple[int, int], x_grid, src: int = 0, gather: bool = True, group=None,
                  eval_tile: Callable | None = None) -> torch.Tensor:
    """Field evaluation of one solved system on a point grid ``x_grid [d, n0, n1, ...]``: the density is broadcast from
    ``src`` (the only collective), then rank r evaluates rows :func:`field_rows` of the grid.  Returns the whole field
    ``[n0, n1, ...]`` when ``gather`` else this rank's rows."""
    rank, world = dist_info(group)
    x_grid = np.asarray(x_grid)
    n0 = x_grid.shape[1]
    dens = broadcast_density(density, density_shape, src=src, group=group)
    sl = field_rows(n0, rank, world)
    tile = x_grid[:, sl]
    fn = eval_tile or _default_eval_tile
    if sl.stop > sl.start:
        u = torch.as_tensor(fn(c, centers, radii, k, eta, n_end, dens, tile)).reshape(tile.shape[1:])
    else:
        # an empty shard lives where the other ranks' tiles live, so that every rank returns a tensor of the same device
        u = torch.zeros((0,) + tuple(x_grid.shape[2:]), dtype=torch.complex128,
                        device=_comm_device(group) if world > 1 else dens.device)
    if not gather or world == 1:
        return u
    return all_gather_rows(u, n0, group)


def ball_rows(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range of row balls [b_lo, b_hi) assembled by ``rank`` (sizes differ by at most one ball)."""
    sl = field_rows(B, rank, world)
    return sl.start, sl.stop


def _default_assemble_rows(c, centers, radii, k, eta, n_end, alpha, beta, b_lo, b_hi):
    from . import _ops
    from ._coords import tree_spec

    spec = tree_spec(c)
    d = spec.d
    centers = np.asarray(centers)[:, list(spec.axes)]  # chain frame
    dev = torch.device("cuda", torch.cuda.current_device())
    kc = complex(k)
    kk = torch.tensor([kc.real], dtype=torch.float64, device=dev)
    kim = torch.tensor([kc.imag], dtype=torch.float64, device=dev) if kc.imag != 0.0 else None
    et = torch.tensor([float(eta)], dtype=torch.float64, device=dev)
    B = np.asarray(radii).shape[0]
    al = torch.as_tensor(np.broadcast_to(np.asarray(alpha, dtype=np.complex128), (B,)).copy(), device=dev)
    be = torch.as_tensor(np.broadcast_to(np.asarray(beta, dtype=np.complex128), (B,)).copy(), device=dev)
    return _ops.assemble(d, n_end, torch.as_tensor(np.asarray(centers), device=dev),
                         torch.as_tensor(np.asarray(radii), device=dev), kk, et, al, be, k_im=kim, rows=(b_lo, b_hi))[0]


def assemble_sharded(c: Any, *, centers, radii, k, eta: float, n_end: int, alpha=1.0, beta=0.0, gather: bool = True,
                     group=None, assemble_rows: Callable | None = None) -> torch.Tensor:
    """System matrix of ONE wavenumber with the block rows split over the ranks (rank r builds the rows of the row balls
    :func:`ball_rows`), then -- if ``gather`` -- an all-gather of the strips so that every rank holds the full
    ``[N, N]`` matrix.  Returns this rank's strip ``[(b_hi - b_lo) H, N]`` otherwise."""
    rank, world = dist_info(group)
    B = np.asarray(radii).shape[0]
    b_lo, b_hi = ball_rows(B, rank, world)
    fn = assemble_rows or _default_assemble_rows
    if b_hi > b_lo:
        strip = torch.as_tensor(fn(c, centers, radii, k, eta, n_end, alpha, beta, b_lo, b_hi))
    else:
        strip = None
    if not gather or world == 1:
        return strip
    meta = [None] * world
    dist.all_gather_object(meta, None if strip is None else (int(strip.shape[0] // (b_hi - b_lo)), int(strip.shape[1])), group=group)
    H, N = next(m for m in meta if m is not None)
    dev = _comm_device(group)
    bmax = -(-B // world)
    buf = torch.zeros((bmax * H, N), dtype=torch.complex128, device=dev)
    if strip is not None:
        buf[: strip.shape[0]] = strip.to(dev)
    parts = [torch.empty_like(torch.view_as_real(buf)) for _ in range(world)]
    dist.all_gather(parts, torch.view_as_real(buf).contiguous(), group=group)
    out = []
    for r in range(world):
        lo, hi = ball_rows(B, r, world)
        out.append(torch.view_as_complex(parts[r])[: (hi - lo) * H])
    full = torch.cat(out, dim=0)
    return full if strip is None else full.to(strip.device)
