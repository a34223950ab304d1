#!/usr/bin/env python3
"""Generate ks265codec_b200/csrc/cuda/ks_dct_gen.cuh: fully unrolled HEVC partial-butterfly transform passes with the
matrix entries as IMMEDIATES (no constant-bank traffic: the 4 KB matrix thrashes the SM's immediate-constant cache
when every IMAD names a different c[bank][offset] operand -- measured: ks_recon_inter_kernel at 10 % issue utilisation).
Exact integer regrouping of  out[u] = (sum_x M[u][x]*in[x] + rnd) >> shift  (no intermediate rounding)."""
import sys

COSV = [64,90,90,90,89,88,87,85,83,82,80,78,75,73,70,67,64,61,57,54,50,46,43,38,36,31,25,22,18,13,9,4,0]


def M(k, n):
    m = (k * (2 * n + 1)) & 127
    return COSV[m] if m <= 32 else -COSV[64 - m] if m <= 64 else -COSV[m - 64] if m <= 96 else COSV[128 - m]


def term(c, v):
    return "" if c == 0 else (" + %d * %s" % (c, v) if c > 0 else " - %d * %s" % (-c, v))


def fwd(N):
    """forward pass, recursive even/odd split.  Interface: in[] registers, store(u, value) functor.
    N == 32: level 0 (the 16x16 odd part, 3/4 of the MACs) runs as a ROLLED loop over rows of a shared-memory table t0
    (t0[j][x] = M32[2j+1][x]) so the instruction footprint stays small; the remaining levels are unrolled immediates."""
    step = 32 // N
    L = ["template <class F> __device__ __forceinline__ void ks_fwd_pass%d(const int (&in)[%d], int shift, const int *__restrict__ t0, F &&store)" % (N, N), "{",
         "    const int rnd = 1 << (shift - 1);"]
    cur, n, lvl, stride = ["in[%d]" % i for i in range(N)], N, 0, 1
    while n >= 2:
        h = n // 2
        e = ["e%d_%d" % (lvl, i) for i in range(h)]; o = ["o%d_%d" % (lvl, i) for i in range(h)]
        for i in range(h):
            L.append("    const int %s = %s + %s, %s = %s - %s;" % (e[i], cur[i], cur[n - 1 - i], o[i], cur[i], cur[n - 1 - i]))
        if N == 32 and lvl == 0:
            L.append("#pragma unroll 2")
            L.append("    for (int j = 0; j < 16; j++) {")
            L.append("        const int4 *t = reinterpret_cast<const int4 *>(t0 + 16 * j);")
            L.append("        const int4 a = t[0], b = t[1], c = t[2], d = t[3];")
            terms = ["a.x", "a.y", "a.z", "a.w", "b.x", "b.y", "b.z", "b.w", "c.x", "c.y", "c.z", "c.w", "d.x", "d.y", "d.z", "d.w"]
            L.append("        int acc = rnd" + "".join(" + %s * %s" % (terms[x], o[x]) for x in range(16)) + ";")
            L.append("        store(2 * j + 1, acc >> shift);")
            L.append("    }")
        else:
            for j in range(h):
                u = stride * (2 * j + 1)
                expr = "".join(term(M(u * step, x), o[x]) for x in range(h))
                L.append("    store(%d, (rnd%s) >> shift);" % (u, expr))
        cur, n, lvl, stride = e, h, lvl + 1, stride * 2
    L.append("    store(0, (rnd + 64 * %s) >> shift);" % cur[0])
    L.append("}")
    return "\n".join(L)


def inv(N):
    """inverse pass: out[y] = sum_k M[k][y]*ld(k).  ld(k) functor fetches input k.  N == 32: the odd inputs (level 0) are
    consumed in a ROLLED loop, accumulators static in registers, coefficients from the same shared-memory table t0."""
    step = 32 // N
    L = ["template <class F> __device__ __forceinline__ void ks_inv_pass%d(F &&ld, int (&out)[%d], int shift, bool clip16, const int *__restrict__ t0)" % (N, N), "{",
         "    const int rnd = 1 << (shift - 1);"]
    used = set()

    def inp(k):
        name = "x%d" % k
        if k not in used:
            used.add(k); L.append("    const int %s = ld(%d);" % (name, k))
        return name

    def rec(n, kstep, tag):
        if n == 1:
            name = "z%s" % tag
            L.append("    const int %s = 64 * %s;" % (name, inp(0)))
            return [name]
        h = n // 2
        ev = rec(h, kstep * 2, tag + "e")
        od = []
        if N == 32 and kstep == 1:
            for y in range(16):
                L.append("    int oL0_%d = 0;" % y); od.append("oL0_%d" % y)
            L.append("#pragma unroll 2")
            L.append("    for (int j = 0; j < 16; j++) {")
            L.append("        const int xi = ld(2 * j + 1);")
            L.append("        const int4 *t = reinterpret_cast<const int4 *>(t0 + 16 * j);")
            L.append("        const int4 a = t[0], b = t[1], c = t[2], d = t[3];")
            terms = ["a.x", "a.y", "a.z", "a.w", "b.x", "b.y", "b.z", "b.w", "c.x", "c.y", "c.z", "c.w", "d.x", "d.y", "d.z", "d.w"]
            for y in range(16):
                L.append("        oL0_%d += %s * xi;" % (y, terms[y]))
            L.append("    }")
        else:
            for y in range(h):
                name = "o%s_%d" % (tag, y)
                expr = "".join(term(M(k * step, y), inp(k)) for k in range(kstep, N, 2 * kstep))
                L.append("    const int %s = 0%s;" % (name, expr))
                od.append(name)
        res = [None] * n
        for y in range(h):
            a, b = "s%s_%d" % (tag, y), "s%s_%d" % (tag, n - 1 - y)
            L.append("    const int %s = %s + %s, %s = %s - %s;" % (a, ev[y], od[y], b, ev[y], od[y]))
            res[y], res[n - 1 - y] = a, b
        return res
    res = rec(N, 1, "")
    for y in range(N):
        L.append("    { int v = (%s + rnd) >> shift; out[%d] = clip16 ? ks_clip3(-32768, 32767, v) : v; }" % (res[y], y))
    L.append("}")
    return "\n".join(L)


def main():
    out = ["/* GENERATED by tools/gen_dct.py -- do not edit.  HEVC core transform passes (H.265 8.6.4.2 matrix) as fully unrolled",
           " * partial butterflies with immediate coefficients.  Reference counterparts: H265_Dct{4,8,16,32}x*_c E@0x4b6230.. (forward,",
           " * partial butterfly) and H265_2dIDct*_c E@0x4417f0.. (inverse). */", "#pragma once", "",
           "/* t0: 16x16 int table in SHARED memory, t0[j][x] = M32[2j+1][x] (only the 32-point passes read it) */", ""]
    for n in (4, 8, 16, 32):
        out.append(fwd(n)); out.append(""); out.append(inv(n)); out.append("")
    out.append("template <int N, class F> __device__ __forceinline__ void ks_fwd_pass(const int (&in)[N], int shift, const int *t0, F &&store)")
    out.append("{ if constexpr (N == 4) ks_fwd_pass4(in, shift, t0, store); else if constexpr (N == 8) ks_fwd_pass8(in, shift, t0, store); else if constexpr (N == 16) ks_fwd_pass16(in, shift, t0, store); else ks_fwd_pass32(in, shift, t0, store); }")
    out.append("template <int N, class F> __device__ __forceinline__ void ks_inv_pass(F &&ld, int (&out)[N], int shift, bool clip16, const int *t0)")
    out.append("{ if constexpr (N == 4) ks_inv_pass4(ld, out, shift, clip16, t0); else if constexpr (N == 8) ks_inv_pass8(ld, out, shift, clip16, t0); else if constexpr (N == 16) ks_inv_pass16(ld, out, shift, clip16, t0); else ks_inv_pass32(ld, out, shift, clip16, t0); }")
    open(sys.argv[1], "w").write("\n".join(out))


if __name__ == "__main__":
    main()
