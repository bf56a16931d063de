"""Comparison of the two assembly kernels (register-resident vs. the legacy shared-memory-resident one, selected per call
by the environment variable BHS_ASM_LEGACY) on a few shapes; prints where and by how much they differ, if at all.

With BHS_ASM_NOPM=1 (no merging of opposite translations) the two are bit-identical; with the merging the blocks of (b', b)
are derived from those of (b, b') by the sign (-1)^(n+n'), which differs from a direct evaluation at -t in the last bits of
the harmonics (expected: relative differences ~1e-15)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops
from biem_helmholtz_sphere_b200.geometry import grid_centers

dev = torch.device("cuda")
torch.manual_seed(0)
bad = 0
for d, half, n_end, nsys, jitter in ((3, 2, 16, 2, 0.0), (3, 2, 16, 1, 0.0), (3, 1, 7, 3, 0.3), (2, 2, 20, 2, 0.0), (3, 2, 24, 1, 0.0),
                                     (4, 1, 4, 2, 0.1), (3, 3, 10, 2, 0.0), (3, 4, 6, 1, 0.0)):
    cen = torch.as_tensor(grid_centers(half, d), device=dev)
    B = cen.shape[0]
    if jitter:
        cen = cen + jitter * torch.rand_like(cen)
    rad = 0.5 + torch.rand(B, dtype=torch.float64, device=dev)
    k = torch.linspace(0.7, 2.9, nsys, dtype=torch.float64, device=dev)
    for rep in range(3):
        os.environ.pop("BHS_ASM_LEGACY", None)
        A = _ops.assemble(d, n_end, cen, rad, k, k)
        os.environ["BHS_ASM_LEGACY"] = "1"
        A0 = _ops.assemble(d, n_end, cen, rad, k, k)
        os.environ.pop("BHS_ASM_LEGACY", None)
        torch.cuda.synchronize()
        same = torch.equal(A.view(torch.float64).view(torch.int64), A0.view(torch.float64).view(torch.int64))
        msg = "identical"
        if not same:
            bad += 1
            diff = (A - A0).abs()
            idx = torch.nonzero(diff > 0)
            if idx.shape[0] == 0:
                print(f"d={d} B={B} n_end={n_end} nsys={nsys} rep {rep}: equal as numbers (signed zeros differ)")
                continue
            H = A.shape[-1] // B
            first = idx[0].tolist()
            rel = (diff / A0.abs().clamp_min(1e-300)).max().item()
            msg = (f"DIFFER: {idx.shape[0]} entries, max |diff| {diff.max().item():.3e} (max |A| {A0.abs().max().item():.3e}), max rel {rel:.2e}, first at sys {first[0]} "
                   f"row (b {first[1] // H}, h {first[1] % H}) col (b' {first[2] // H}, h' {first[2] % H}); rows hit {idx[:,1].unique().numel()}, cols hit {idx[:,2].unique().numel()}")
        print(f"d={d} B={B} n_end={n_end} nsys={nsys} rep {rep}: {msg}")
print("ALL IDENTICAL" if bad == 0 else f"{bad} MISMATCHES")
