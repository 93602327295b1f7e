"""ORACLE (test infrastructure, NOT product code): BVH light sampler construction restated from the reference.

Follows src/lights/bvh-light-sampler.jl:237-466 (`_evaluate_cost`, `BVHLightSampler(lights)`, `_build_bvh!`) and
src/lights/light-bounds.jl:21-158, 231-295 (DirectionCone / LightBounds unions, `light_bounds` per light type) with
Float32 scalars and the reference's 1-based indices, written independently of csrc/host_lightbvh.cpp (the product-side
builder it checks, tests/test_host_logic.py::test_light_bvh_builder_matches_restatement).  Pure-Python loops: meant for a
few hundred lights.  Transcendentals are evaluated in double and rounded to Float32 (Julia's Float32 acos / asin / sin /
cos are within an ulp of that), so float fields are compared to 1e-5 relative while the tree STRUCTURE (child indices,
leaf flags, light ids, bit trails, infinite-light list) must be identical.
"""
import math

import numpy as np

f32 = np.float32
PI = f32(math.pi)
INF = f32(np.inf)
NUM_BUCKETS = 12          # bvh-light-sampler.jl:237


def _f(x):
    return f32(x)


def _acos(x): return f32(math.acos(float(min(max(x, f32(-1)), f32(1)))))
def _asin(x): return f32(math.asin(float(min(max(x, f32(-1)), f32(1)))))
def _sin(x): return f32(math.sin(float(x)))
def _cos(x): return f32(math.cos(float(x)))


def _v(x, y, z):
    return np.array([x, y, z], dtype=f32)


def _dot(a, b):
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def _cross(a, b):
    return _v(f32(a[1] * b[2]) - f32(a[2] * b[1]), f32(a[2] * b[0]) - f32(a[0] * b[2]), f32(a[0] * b[1]) - f32(a[1] * b[0]))


def _norm(a):
    return f32(np.sqrt(_dot(a, a)))


def _normalize(a):
    inv = f32(1) / _norm(a)
    return _v(inv * a[0], inv * a[1], inv * a[2])


class LightBounds:                      # light-bounds.jl:104-113
    def __init__(self, lo=None, hi=None, w=None, phi=0.0, cos_o=1.0, cos_e=1.0, two_sided=False):
        self.lo = _v(INF, INF, INF) if lo is None else np.asarray(lo, dtype=f32)
        self.hi = _v(-INF, -INF, -INF) if hi is None else np.asarray(hi, dtype=f32)
        self.w = _v(0, 0, 1) if w is None else np.asarray(w, dtype=f32)
        self.phi, self.cos_o, self.cos_e, self.two_sided = f32(phi), f32(cos_o), f32(cos_e), bool(two_sided)

    def centroid(self):                 # :115
        return (self.lo + self.hi) * f32(0.5)


def angle_between(a, b):                # light-bounds.jl:44-50
    if _dot(a, b) < 0:
        return PI - f32(2) * _asin(_norm(a + b) * f32(0.5))
    return f32(2) * _asin(_norm(b - a) * f32(0.5))


def cone_union(aw, ac, bw, bc):         # light-bounds.jl:58-87; returns (w, cos)
    if ac == INF:
        return bw, bc
    if bc == INF:
        return aw, ac
    ta, tb = _acos(ac), _acos(bc)
    td = angle_between(aw, bw)
    if min(td + tb, PI) <= ta:
        return aw, ac
    if min(td + ta, PI) <= tb:
        return bw, bc
    to = f32(f32(f32(ta + td) + tb) * f32(0.5))
    if to >= PI:
        return _v(0, 0, 1), f32(-1)
    tr = f32(to - ta)
    wr = _cross(aw, bw)
    if _dot(wr, wr) == 0:
        return _v(0, 0, 1), f32(-1)
    ax = _normalize(wr)
    s, c = _sin(tr), _cos(tr)
    w = aw * c + _cross(ax, aw) * s + ax * _dot(ax, aw) * f32(f32(1) - c)
    return _normalize(w.astype(f32)), _cos(to)


def lb_union(a, b):                     # light-bounds.jl:142-158
    if a.phi == 0:
        return b
    if b.phi == 0:
        return a
    w, c = cone_union(a.w, a.cos_o, b.w, b.cos_o)
    return LightBounds(np.minimum(a.lo, b.lo), np.maximum(a.hi, b.hi), w, f32(a.phi + b.phi), c, min(a.cos_e, b.cos_e), a.two_sided or b.two_sided)


# ---- light_bounds per light type (light-bounds.jl:231-295), from the flattened HkLight records --------------------------
HK_LIGHT_POINT, HK_LIGHT_SPOT, HK_LIGHT_DIFFUSE_AREA = 1, 2, 7


def _sigmoid(x):
    if math.isinf(float(x)):
        return f32(1) if x > 0 else f32(0)
    return f32(f32(0.5) + x / f32(f32(2) * f32(np.sqrt(f32(f32(1) + f32(x * x))))))


def _poly(p, l):
    return _sigmoid(f32(f32(f32(f32(p[0] * l) * l) + f32(p[1] * l)) + p[2]))


def _luminance(L):                      # light-sampler.jl:444-456 (RGBIlluminantSpectrum: scale * max value * 100; RGB: Y)
    if L.spectrum_kind == 1:
        p = [f32(v) for v in L.poly]
        r = max(_poly(p, f32(360)), _poly(p, f32(830)))
        if p[0] != 0:
            lc = f32(-p[1] / f32(f32(2) * p[0]))
            if f32(360) <= lc <= f32(830):
                r = max(r, _poly(p, lc))
        return f32(f32(f32(L.illum_scale) * r) * f32(100))
    return f32(f32(f32(f32(0.212671) * f32(L.rgb[0])) + f32(f32(0.715160) * f32(L.rgb[1]))) + f32(f32(0.072169) * f32(L.rgb[2])))


def light_bounds(L):
    cos_pi, cos_half_pi = f32(math.cos(math.pi)), f32(math.cos(math.pi / 2))
    if L.type == HK_LIGHT_POINT:        # :231-245
        p = _v(*L.position)
        return LightBounds(p, p, _v(0, 0, 1), f32(f32(f32(f32(4) * PI) * f32(L.scale)) * _luminance(L)), cos_pi, cos_half_pi, False)
    if L.type == HK_LIGHT_SPOT:         # :248-270; light_to_world * (0,0,1) = third row of the rigid world_to_light
        p = _v(*L.position)
        w = _normalize(_v(L.world_to_light[8], L.world_to_light[9], L.world_to_light[10]))
        ce = f32(math.cos(float(f32(_acos(f32(L.cos_total_width)) - _acos(f32(L.cos_falloff_start))))))
        if ce == 1 and L.cos_total_width != L.cos_falloff_start:
            ce = f32(0.999)
        return LightBounds(p, p, w, f32(f32(f32(f32(4) * PI) * f32(L.scale)) * _luminance(L)), f32(L.cos_falloff_start), ce, False)
    if L.type == HK_LIGHT_DIFFUSE_AREA:  # :273-295
        v = np.array(list(L.v), dtype=f32).reshape(3, 3)
        lum = f32(f32(f32(f32(0.212671) * f32(L.rgb[0])) + f32(f32(0.715160) * f32(L.rgb[1]))) + f32(f32(0.072169) * f32(L.rgb[2])))
        sided = f32(2) if L.two_sided else f32(1)
        phi = f32(f32(f32(f32(PI * sided) * f32(L.area)) * f32(L.scale)) * lum)
        return LightBounds(v.min(axis=0), v.max(axis=0), _v(*L.normal), phi, f32(1), cos_half_pi, bool(L.two_sided))
    return None


def evaluate_cost(lb, lo, hi, dim):     # bvh-light-sampler.jl:242-258 (dim 0-based here)
    to, te = _acos(lb.cos_o), _acos(lb.cos_e)
    tw = min(f32(to + te), PI)
    so = f32(np.sqrt(max(f32(0), f32(f32(1) - f32(lb.cos_o * lb.cos_o)))))
    inner = f32(f32(f32(f32(f32(f32(2) * tw) * so) - _cos(f32(to - f32(f32(2) * tw)))) - f32(f32(f32(2) * to) * so)) + lb.cos_o)
    M = f32(f32(f32(f32(2) * PI) * f32(f32(1) - lb.cos_o)) + f32(f32(PI / f32(2)) * inner))
    d = (hi - lo).astype(f32)
    md = max(d[0], d[1], d[2])
    Kr = f32(md / d[dim]) if d[dim] > f32(1e-10) else f32(md / f32(1e-10))
    sa = f32(f32(2) * f32(f32(f32(d[0] * d[1]) + f32(d[0] * d[2])) + f32(d[1] * d[2])))
    return f32(f32(f32(lb.phi * M) * Kr) * sa)


def _bucket(cb_lo, cb_hi, c, dim):      # floor(12 * Raycore.offset(bounds, c)[dim]), clamped, 1-based
    o = f32(c[dim] - cb_lo[dim])
    if cb_hi[dim] > cb_lo[dim]:
        o = f32(o / f32(cb_hi[dim] - cb_lo[dim]))
    b = int(math.floor(float(f32(f32(NUM_BUCKETS) * o))))
    return min(max(b, 0), NUM_BUCKETS - 1) + 1


def build_light_sampler(lights):
    """lights: sequence of HkLight.  Returns dict(nodes=[...], trails=uint32[n], infinite=[1-based], n_bvh)."""
    n = len(lights)
    items, infinite = [], []
    for flat in range(1, n + 1):
        lb = light_bounds(lights[flat - 1])
        if lb is None:
            infinite.append(flat)
        elif lb.phi > 0:
            items.append((flat, lb))
    trails = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    nodes = []

    def node(lb, child, leaf):
        return dict(lo=lb.lo.copy(), hi=lb.hi.copy(), w=lb.w.copy(), phi=lb.phi, cos_o=lb.cos_o, cos_e=lb.cos_e, two_sided=lb.two_sided, child=child, leaf=leaf)

    def build(start, stop, trail, depth):   # 1-based inclusive, :337-466
        count = stop - start + 1
        if count == 1:
            flat, lb = items[start - 1]
            nodes.append(node(lb, flat, True))
            trails[flat - 1] = trail
            return lb
        overall = items[start - 1][1]
        c = overall.centroid()
        cb_lo, cb_hi = c.copy(), c.copy()
        for i in range(start + 1, stop + 1):
            overall = lb_union(overall, items[i - 1][1])
            c = items[i - 1][1].centroid()
            cb_lo, cb_hi = np.minimum(cb_lo, c), np.maximum(cb_hi, c)
        best_cost, best_dim, best_bucket = INF, -1, 0
        for dim in range(3):
            if cb_hi[dim] - cb_lo[dim] <= 0:
                continue
            bb = [LightBounds() for _ in range(NUM_BUCKETS)]
            bc = [0] * NUM_BUCKETS
            for i in range(start, stop + 1):
                b = _bucket(cb_lo, cb_hi, items[i - 1][1].centroid(), dim)
                bb[b - 1] = lb_union(bb[b - 1], items[i - 1][1]); bc[b - 1] += 1
            for split in range(1, NUM_BUCKETS):
                below, above, nb, na = LightBounds(), LightBounds(), 0, 0
                for b in range(1, split + 1):
                    below = lb_union(below, bb[b - 1]); nb += bc[b - 1]
                for b in range(split + 1, NUM_BUCKETS + 1):
                    above = lb_union(above, bb[b - 1]); na += bc[b - 1]
                if nb == 0 or na == 0:
                    continue
                cost = f32(evaluate_cost(below, overall.lo, overall.hi, dim) + evaluate_cost(above, overall.lo, overall.hi, dim))
                if cost < best_cost:
                    best_cost, best_dim, best_bucket = cost, dim, split
        if best_dim >= 0:
            pivot = start
            for i in range(start, stop + 1):
                if _bucket(cb_lo, cb_hi, items[i - 1][1].centroid(), best_dim) <= best_bucket:
                    if i != pivot:
                        items[pivot - 1], items[i - 1] = items[i - 1], items[pivot - 1]
                    pivot += 1
            mid = start + count // 2 if (pivot == start or pivot > stop) else pivot - 1
        else:
            mid = start + count // 2 - 1
        mid = min(max(mid, start), stop - 1)
        me = len(nodes)
        nodes.append(node(overall, 0, False))
        l0 = build(start, mid, trail, depth + 1)
        child1 = len(nodes) + 1
        l1 = build(mid + 1, stop, trail | (1 << depth), depth + 1)
        merged = lb_union(l0, l1)
        nodes[me] = node(merged, child1, False)
        return merged

    if items:
        build(1, len(items), 0, 0)
    return dict(nodes=nodes, trails=trails, infinite=infinite, n_bvh=len(items))
