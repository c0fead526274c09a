"""Synthetic Legion datasets: degree-skewed directed CSR
graphs with closed-form float32 features, defined by pure 64-bit integer
arithmetic so the numpy (CPU) and torch (GPU) builds are bit-identical.

Shapes follow the launcher's dataset table (legion_server.py:6-53):
indptr int64[N+1], indices int32[E], features float32[N, D], labels int32[N],
train/valid/test id lists (GPUGraphStore.cu:266-301).
"""
import numpy as np

_M64 = (1 << 64) - 1
SEED = 0x1E6104


class _NP:
    """uint64 helpers on numpy arrays."""
    name = "numpy"

    @staticmethod
    def arange(a, b, device=None):
        return np.arange(a, b, dtype=np.uint64)

    @staticmethod
    def mul(x, c):
        return x * np.uint64(c & _M64)

    @staticmethod
    def add(x, c):
        return x + np.uint64(c & _M64)

    @staticmethod
    def lsr(x, s):
        return x >> np.uint64(s)

    @staticmethod
    def band(x, m):
        return x & np.uint64(m)

    @staticmethod
    def to_i64(x):
        return x.astype(np.int64)


class _TH:
    """the same on torch int64 tensors (two's-complement wrap-around)."""
    name = "torch"

    @staticmethod
    def _c(c):
        c &= _M64
        return c - (1 << 64) if c >= (1 << 63) else c

    @staticmethod
    def arange(a, b, device=None):
        import torch
        return torch.arange(a, b, dtype=torch.int64, device=device)

    @classmethod
    def mul(cls, x, c):
        return x * cls._c(c)

    @classmethod
    def add(cls, x, c):
        return x + cls._c(c)

    @staticmethod
    def lsr(x, s):
        return (x >> s) & ((1 << (64 - s)) - 1)

    @classmethod
    def band(cls, x, m):
        return x & cls._c(m)

    @staticmethod
    def to_i64(x):
        return x


def _mix(xp, x):
    """splitmix64 finaliser."""
    x = xp.add(x, 0x9E3779B97F4A7C15)
    x = xp.mul(x ^ xp.lsr(x, 30), 0xBF58476D1CE4E5B9)
    x = xp.mul(x ^ xp.lsr(x, 27), 0x94D049BB133111EB)
    return x ^ xp.lsr(x, 31)


def _degrees(xp, lo, hi, n_nodes, dmin_fp, kmax, max_deg, seed, device=None):
    """Discrete power law: deg doubles with halving probability (alpha = 2),
    linear interpolation inside an octave, 16.16 fixed point."""
    i = xp.arange(lo, hi, device)
    h = _mix(xp, xp.add(xp.mul(i, 0xD6E8FEB86659FD93), seed))
    u = xp.to_i64(xp.band(h, 0xFFFFFFFF))
    frac = xp.to_i64(xp.band(xp.lsr(h, 32), 0xFFFF))
    z = u * 0
    for k in range(1, kmax + 1):
        z = z + (u < (1 << (32 - k)))
    base = (z * 0 + dmin_fp) << z
    deg = (base + ((base * frac) >> 16)) >> 16
    lim = min(max_deg, n_nodes - 1)
    if xp is _TH:
        return deg.clamp(max=lim)
    return np.minimum(deg, lim)


def _neighbours(xp, lo, hi, n_nodes, scatter, seed, device=None):
    """indices[e] for global edge slots e in [lo, hi): product of three uniforms
    (skewed towards 0) scattered over the id range by a multiplicative bijection."""
    e = xp.arange(lo, hi, device)
    h = _mix(xp, xp.add(xp.mul(e, 0xA24BAED4963EE407), seed ^ 0x5851F42D4C957F2D))
    u1 = xp.band(h, 0xFFFFFFFF)
    u2 = xp.lsr(h, 32)
    u3 = xp.band(_mix(xp, h), 0xFFFFFFFF)
    x = xp.lsr(xp.lsr(u1 * u2, 32) * u3, 32)          # 32-bit fixed point in [0,1)
    r = xp.lsr(xp.mul(x, n_nodes), 32)                # rank in [0, N)
    r = xp.to_i64(r)
    dst = (r * scatter) % n_nodes
    return dst


def _scatter_mult(n_nodes):
    """odd multiplier < 2^31 co-prime with N so r -> r*m mod N is a bijection and r*m < 2^63."""
    from math import gcd
    m = 0x9E3779B1 % n_nodes or 1
    m = min(m, (1 << 31) - 1)
    while gcd(m, n_nodes) != 1:
        m += 1
    return m


def calibrate_dmin(avg_deg, n_nodes, kmax=6, max_deg=10000, seed=SEED, sample=200_000):
    """bisection on the 16.16 fixed-point base degree so the mean degree of a
    deterministic sample of nodes matches avg_deg."""
    s = min(sample, n_nodes)
    lo, hi = 1, int(avg_deg * 65536 * 4) + 65536
    while lo < hi:
        mid = (lo + hi) // 2
        m = float(_degrees(_NP, 0, s, n_nodes, mid, kmax, max_deg, seed).mean())
        if m < avg_deg:
            lo = mid + 1
        else:
            hi = mid
    return lo


class Dataset:
    """Container: arrays are numpy (backend='numpy') or torch tensors on `device`."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def features_block(xp, lo, hi, dim, device=None):
    """feat[i][j] = bit_cast<float>(0x3F000000 + ((i*2654435761 + j*40503) & 0x7FFFFF)) in [0.5, 1)."""
    i = xp.arange(lo, hi, device)
    j = xp.arange(0, dim, device)
    if xp is _TH:
        import torch
        v = (i[:, None] * 2654435761 + j[None, :] * 40503) & 0x7FFFFF
        return (v + 0x3F000000).to(torch.int32).view(torch.float32)
    v = (i[:, None] * np.uint64(2654435761) + j[None, :] * np.uint64(40503)) & np.uint64(0x7FFFFF)
    return (v + np.uint64(0x3F000000)).astype(np.uint32).view(np.float32)


def make_dataset(n_nodes, avg_deg, dim, n_class=47, backend="numpy", device=None, seed=SEED,
                 kmax=6, max_deg=10000, with_features=True, chunk=1 << 24, dmin_fp=None):
    xp = _NP if backend == "numpy" else _TH
    if dmin_fp is None:
        dmin_fp = calibrate_dmin(avg_deg, n_nodes, kmax, max_deg, seed)
    scatter = _scatter_mult(n_nodes)
    if xp is _TH:
        import torch
        deg = torch.cat([_degrees(xp, lo, min(lo + chunk, n_nodes), n_nodes, dmin_fp, kmax, max_deg, seed, device)
                         for lo in range(0, n_nodes, chunk)])
        indptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=device)
        torch.cumsum(deg, 0, out=indptr[1:])
        n_edges = int(indptr[-1].item())
        indices = torch.empty(n_edges, dtype=torch.int32, device=device)
        for lo in range(0, n_edges, chunk):
            hi = min(lo + chunk, n_edges)
            indices[lo:hi] = _neighbours(xp, lo, hi, n_nodes, scatter, seed, device).to(torch.int32)
        feats = None
        if with_features:
            feats = torch.empty((n_nodes, dim), dtype=torch.float32, device=device)
            rows = max(1, chunk // dim)
            for lo in range(0, n_nodes, rows):
                hi = min(lo + rows, n_nodes)
                feats[lo:hi] = features_block(xp, lo, hi, dim, device)
        labels = (torch.arange(n_nodes, dtype=torch.int64, device=device) % n_class).to(torch.int32)
        ids = torch.arange(n_nodes, dtype=torch.int32, device=device)
    else:
        deg = _degrees(xp, 0, n_nodes, n_nodes, dmin_fp, kmax, max_deg, seed)
        indptr = np.zeros(n_nodes + 1, np.int64)
        np.cumsum(deg, out=indptr[1:])
        n_edges = int(indptr[-1])
        indices = np.empty(n_edges, np.int32)
        for lo in range(0, n_edges, chunk):
            hi = min(lo + chunk, n_edges)
            indices[lo:hi] = _neighbours(xp, lo, hi, n_nodes, scatter, seed).astype(np.int32)
        feats = None
        if with_features:
            feats = np.empty((n_nodes, dim), np.float32)
            rows = max(1, chunk // dim)
            for lo in range(0, n_nodes, rows):
                hi = min(lo + rows, n_nodes)
                feats[lo:hi] = features_block(xp, lo, hi, dim)
        labels = (np.arange(n_nodes, dtype=np.int64) % n_class).astype(np.int32)
        ids = np.arange(n_nodes, dtype=np.int32)
    # 10 % train split (legion_server.py:19,35), 1 % valid, 1 % test: a multiplicative hash of the id
    # picks the subsets so that `tid % P` partitions (GPUGraphStore.cu:338) stay balanced; file order = id order
    key = (ids.to(torch.int64) if xp is _TH else ids.astype(np.int64)) * 2654435761 & 0xFFFFFFFF
    t1, t2, t3 = 429496730, 429496730 + 42949673, 429496730 + 2 * 42949673
    train = ids[key < t1]
    valid = ids[(key >= t1) & (key < t2)]
    test = ids[(key >= t2) & (key < t3)]
    return Dataset(n_nodes=n_nodes, n_edges=n_edges, dim=dim, n_class=n_class, indptr=indptr, indices=indices,
                   features=feats, labels=labels, train_ids=train, valid_ids=valid, test_ids=test,
                   dmin_fp=dmin_fp, seed=seed, backend=backend)


def partition_ids(ids, parts):
    """GPUGraphStore.cu:332-346: seed tid goes to partition tid % P (file order kept)."""
    return [ids[(ids % parts) == p] for p in range(parts)]


# dataset shapes of BASELINE.json configs (name -> N, avg_deg, D, classes, batch, fanout)
CONFIGS = {
    "C1": dict(n_nodes=100_000, avg_deg=15.0, dim=128, n_class=47, batch=1024, fanout=[25, 10]),
    "C2": dict(n_nodes=2_449_029, avg_deg=61_859_140 / 2_449_029, dim=100, n_class=47, batch=8000, fanout=[25, 10]),
    "C3": dict(n_nodes=111_059_956, avg_deg=1_615_685_872 / 111_059_956, dim=128, n_class=172, batch=8000,
               fanout=[25, 10]),
    "C4": dict(n_nodes=133_633_040, avg_deg=5_507_679_822 / 133_633_040, dim=256, n_class=2, batch=8000,
               fanout=[25, 10]),
    "C5": dict(n_nodes=65_608_366, avg_deg=1_806_067_135 / 65_608_366, dim=256, n_class=2, batch=8000,
               fanout=[15, 10, 5]),
}
