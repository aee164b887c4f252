import numpy as np
import torch


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def bits_equal(a, b):
    """Bitwise equality, except that -0.0 and +0.0 are allowed to differ only if asked."""
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def mismatch_report(a, b, name=""):
    a = np.asarray(a); b = np.asarray(b)
    bad = a.view(np.uint32) != b.view(np.uint32)
    n = int(bad.sum())
    if n == 0:
        return f"{name}: identical"
    idx = np.argwhere(bad)
    first = tuple(idx[0])
    rows = (int(idx[:, 0].min()), int(idx[:, 0].max())); cols = (int(idx[:, 1].min()), int(idx[:, 1].max()))
    return (f"{name}: {n} cells differ (rows {rows}, cols {cols}); first at {first}: got {a[first]!r} want {b[first]!r}; "
            f"max abs diff {float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64))))}")


def rel_l2(a, b, floor=1e-30):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), floor))
