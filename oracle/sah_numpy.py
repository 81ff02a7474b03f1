"""ORACLE -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement of the reference's exact SAH builder
(`BVHAggregate::create_sah_node`, src/aggregate/bvh.rs:24-124), written in numpy float32 directly
from the Rust and sharing nothing with oracle/oracle.cpp: the tests hold the two restatements
(and through them the product's host and device builders) against each other.  Triangles only:
`Triangle::bound` (src/geometry/shape/triangle.rs:114-119) is a min / max over three vertices and
needs no transform arithmetic.

Arithmetic follows the reference operation for operation, every step rounded to f32 (Rust does
not contract a*b+c):
  Bounds::center        min + (max - min) / 2                         bounds.rs:67-72
  Bounds::surface_area  2 * ((dx*dy + dx*dz) + dy*dz)                 bounds.rs:255-258
  union                 per-component min / max                       bounds.rs:186-199
  ordering              stable sort of the range by centre[axis]      bvh.rs:51-54
                        (`sort_by_cached_key` keeps equal keys in place; OrderedFloat: -0 == +0)
  cost                  0.125 + ((i+1)*SA(fwd_i) + (len-1-i)*SA(bwd_i)) / SA(full)   bvh.rs:72-76
  choice                first strict minimum over axis 0,1,2 then i ascending; the very first
                        candidate is taken unconditionally, even when it is NaN     bvh.rs:77-90
  leaf                  len <= 1, or min cost > len                   bvh.rs:31-36,93-100
  reorder               the cycle-following swap loop, literally      bvh.rs:101-113
  node numbering        post-order: left subtree, right subtree, parent   bvh.rs:114-123
"""
import numpy as np

F = np.float32


def triangle_bounds(tris):
    """tris: (n, 3, 3) float32 vertices -> (lo, hi), each (n, 3) float32."""
    tris = np.asarray(tris, F)
    return tris.min(axis=1), tris.max(axis=1)


def _surface_area(lo, hi):
    d = (hi - lo).astype(F)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return (F(2) * (((dx * dy).astype(F) + (dx * dz).astype(F)).astype(F) + (dy * dz).astype(F)).astype(F)).astype(F)


def _centre(lo, hi):
    return (lo + ((hi - lo).astype(F) / F(2)).astype(F)).astype(F)


class SahTree:
    """Result of build(): `order[k]` = input index of the primitive at position k; nodes in
    creation order as parallel lists (is_leaf, axis, a, b, bound) -- a, b are the primitive range
    of a leaf or the child node indices of a parent; `root` is the last node."""

    def __init__(self, n):
        self.order = np.arange(n, dtype=np.int64)
        self.is_leaf, self.axis, self.a, self.b, self.bound = [], [], [], [], []

    @property
    def root(self):
        return len(self.a) - 1

    def _push(self, is_leaf, axis, a, b, bound):
        self.is_leaf.append(is_leaf); self.axis.append(axis); self.a.append(a); self.b.append(b)
        self.bound.append(bound)
        return len(self.a) - 1


def build(lo, hi):
    """Exact SAH build over primitive bounds lo, hi ((n, 3) float32, n >= 1)."""
    import sys
    lo = np.array(lo, F); hi = np.array(hi, F)
    n = lo.shape[0]
    assert n > 0
    tree = SahTree(n)
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * n + 1000))
    with np.errstate(all="ignore"):
        _node(tree, lo, hi, 0, n)
    return tree


def _node(tree, lo, hi, start, end):
    length = end - start
    if length <= 1:
        return tree._push(1, 0, start, end, np.zeros(6, F))
    plo, phi = lo[start:end], hi[start:end]
    full_lo, full_hi = plo.min(axis=0), phi.max(axis=0)
    full_sa = _surface_area(full_lo, full_hi)
    centre = _centre(plo, phi)
    counts_l = np.arange(1, length, dtype=np.int64).astype(F)          # (i + 1) as Float
    counts_r = (length - np.arange(1, length, dtype=np.int64)).astype(F)   # (len - 1 - i) as Float
    costs, sorts = [], []
    for axis in range(3):
        key = centre[:, axis] + F(0)                                   # -0.0 -> +0.0: OrderedFloat compares them equal
        idx = np.argsort(key, kind="stable")
        fwd_lo = np.minimum.accumulate(plo[idx], axis=0)[:-1]
        fwd_hi = np.maximum.accumulate(phi[idx], axis=0)[:-1]
        bwd_lo = np.minimum.accumulate(plo[idx][::-1], axis=0)[:-1][::-1]
        bwd_hi = np.maximum.accumulate(phi[idx][::-1], axis=0)[:-1][::-1]
        num = ((counts_l * _surface_area(fwd_lo, fwd_hi)).astype(F) + (counts_r * _surface_area(bwd_lo, bwd_hi)).astype(F)).astype(F)
        costs.append((F(0.125) + (num / full_sa).astype(F)).astype(F))
        sorts.append(idx)
    flat = np.concatenate(costs)
    if np.isnan(flat[0]):
        pick = 0                                                       # nothing compares below NaN
    else:
        pick = int(np.argmin(np.where(np.isnan(flat), F(np.inf), flat)))   # first strict minimum
    min_cost, min_axis, min_i = flat[pick], pick // (length - 1), pick % (length - 1)
    if min_cost > F(length):
        return tree._push(1, 0, start, end, np.zeros(6, F))
    s = sorts[min_axis]
    # bvh.rs:101-113, literally; lo / hi / order play the part of self.primitives
    swapped = [False] * length
    for i in range(length):
        if not swapped[i]:
            j = i
            k = int(s[j])
            while k != i:
                for arr in (lo, hi, tree.order):
                    tmp = arr[j + start].copy(); arr[j + start] = arr[k + start]; arr[k + start] = tmp
                swapped[k] = True
                j = k
                k = int(s[k])
    left = _node(tree, lo, hi, start, start + min_i + 1)
    right = _node(tree, lo, hi, start + min_i + 1, end)
    return tree._push(0, min_axis, left, right, np.concatenate([full_lo, full_hi]).astype(F))
