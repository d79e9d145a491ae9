// sched16.cuh -- conflict-free shared-memory schedule for one half-warp of the LDS.64 assignment
// kernel (layout mode 1, see convert.cu).  Host/device code: the device kernel runs it with one
// thread per half-warp on interleaved shared memory, tests/test_sched16.py compiles it with g++
// and checks the colouring on the CPU.
//
// Problem: 16 lanes (columns), lane l holds cnt[l][g] entries whose table row is g mod 16.  The
// centroid table is staged twice; an entry of group g can be read from copy A (bank class g) or
// copy B (class g+1 mod 16).  A step (one LDS.64 per lane) costs one wavefront iff the 16 classes
// touched are distinct.  We need a W-step schedule.
//   1. Choose copies so that no class holds more than W entries: excess flows round the ring
//      c -> c+1 (only group-c entries can move from class c to c+1).
//   2. Edge-colour the bipartite multigraph lanes x classes with W colours (colour = step).
//      Koenig's theorem: possible when every degree is <= W.  Insert edges one at a time; if lane
//      and class share no free colour, take a free at the lane, b free at the class, swap a<->b on
//      the alternating path starting at the class, then use a.
//   3. Steps where a lane has no entry are pads served from a zero row of a class still free.
// Output: code[l * W + t] = group | copy << 4 for an entry, 0x80 | class for a pad.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SKM_HD __host__ __device__
#else
#define SKM_HD
#endif

SKM_HD static inline int skm_ffs32(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)m);
#else
    return __builtin_ffs((int)m);
#endif
}

// NL = lanes per problem = bank classes: 16 (half-warp, 8-byte gathers) or 8 (quarter-warp, 16-byte gathers)
// bytes / words of scratch per problem
SKM_HD static inline int skm_sched_bytes(int nl, int wmax) { return 3 * nl * nl + 2 * nl * wmax; }
SKM_HD static inline int skm_sched_words(int nl, int wmax) { return 2 * nl * ((wmax + 31) >> 5); }
SKM_HD static inline int skm_sched16_bytes(int wmax) { return skm_sched_bytes(16, wmax); }
SKM_HD static inline int skm_sched16_words(int wmax) { return skm_sched_words(16, wmax); }

// Mem: B(i) -> unsigned char&, Wd(i) -> uint32_t&  (scratch of skm_sched16_bytes / _words)
// On entry B(l*NL+g), l,g in [0,NL), holds cnt[l][g]; everything else is scratch.
// Returns the number of entries that could not be scheduled conflict-free (0 when balanced).
template <int NL, class Mem, class Out>
SKM_HD static int skm_sched(Mem &M, int W, int wmax, Out &out)
{
    constexpr int MK = NL - 1;
    const int nw = (wmax + 31) >> 5;
    const int OFF_CNT = 0, OFF_MOV = NL * NL, OFF_OVF = 2 * NL * NL, OFF_AT = 3 * NL * NL, OFF_ATC = 3 * NL * NL + NL * wmax;
    const int OFF_FL = 0, OFF_FC = NL * nw;
    for (int i = NL * NL; i < 3 * NL * NL; ++i) M.B(i) = 0;

    // ---- 1. balance ----
    int T[NL], d[NL], x[NL];
    for (int c = 0; c < NL; ++c) {
        int s = 0;
        for (int l = 0; l < NL; ++l) s += M.B(OFF_CNT + l * NL + c);
        T[c] = s; d[c] = s; x[c] = 0;
    }
    for (int pass = 0; pass < 64; ++pass) {
        bool moved = false;
        for (int c = 0; c < NL; ++c) {
            const int ex = d[c] - W, avail = T[c] - x[c];
            if (ex > 0 && avail > 0) {
                const int mv = ex < avail ? ex : avail;
                x[c] += mv; d[c] -= mv; d[(c + 1) & MK] += mv;
                moved = true;
            }
        }
        if (!moved) break;
    }
    for (int c = 0; c < NL; ++c) {          // spread the moves of group c over the lanes, round-robin
        int rem = x[c];
        while (rem > 0) {
            bool any = false;
            for (int l = 0; l < NL && rem > 0; ++l) {
                const int have = M.B(OFF_CNT + l * NL + c), mv = M.B(OFF_MOV + l * NL + c);
                if (mv < have) { M.B(OFF_MOV + l * NL + c) = (unsigned char)(mv + 1); --rem; any = true; }
            }
            if (!any) break;
        }
    }

    // ---- 2. edge colouring ----
    for (int i = 0; i < 2 * NL * wmax; ++i) M.B(OFF_AT + i) = 0xFF;
    for (int v = 0; v < 2 * NL; ++v)
        for (int q = 0; q < nw; ++q) {
            const int lo = q * 32;
            M.Wd(v * nw + q) = (W - lo >= 32) ? 0xffffffffu : (W > lo ? ((1u << (W - lo)) - 1u) : 0u);
        }
    auto first_free = [&](int off, int v) -> int {
        for (int q = 0; q < nw; ++q) { const uint32_t m = M.Wd(off + v * nw + q); if (m) return q * 32 + skm_ffs32(m) - 1; }
        return -1;
    };
    auto set_free = [&](int off, int v, int t, bool fr) {
        uint32_t m = M.Wd(off + v * nw + (t >> 5));
        if (fr) m |= 1u << (t & 31); else m &= ~(1u << (t & 31));
        M.Wd(off + v * nw + (t >> 5)) = m;
    };
    auto put = [&](int l, int c, int t) {
        M.B(OFF_AT + l * wmax + t) = (unsigned char)c;
        M.B(OFF_ATC + c * wmax + t) = (unsigned char)l;
        set_free(OFF_FL, l, t, false);
        set_free(OFF_FC, c, t, false);
    };
    int overflow = 0;
    for (int l = 0; l < NL; ++l) {
        for (int c = 0; c < NL; ++c) {
            int mult = (int)M.B(OFF_CNT + l * NL + c) - (int)M.B(OFF_MOV + l * NL + c) +
                       (int)M.B(OFF_MOV + l * NL + ((c + MK) & MK));
            for (; mult > 0; --mult) {
                int tc = -1;
                for (int q = 0; q < nw; ++q) {
                    const uint32_t m = M.Wd(OFF_FL + l * nw + q) & M.Wd(OFF_FC + c * nw + q);
                    if (m) { tc = q * 32 + skm_ffs32(m) - 1; break; }
                }
                if (tc >= 0) { put(l, c, tc); continue; }
                const int a = first_free(OFF_FL, l), b = first_free(OFF_FC, c);
                if (a < 0) { ++overflow; continue; }                         // lane longer than W: caller's bug
                if (b < 0) { M.B(OFF_OVF + l * NL + c) += 1; ++overflow; continue; }   // class holds more than W entries
                int pl[34], pc[34], np = 0;         // path edges (lane, class), colours a, b, a, ...
                int vc = c;
                while (np < 2 * NL) {
                    const int l1 = M.B(OFF_ATC + vc * wmax + a);
                    if (l1 == 0xFF) break;
                    pl[np] = l1; pc[np] = vc; ++np;
                    const int c2 = M.B(OFF_AT + l1 * wmax + b);
                    if (c2 == 0xFF) break;
                    pl[np] = l1; pc[np] = c2; ++np;
                    vc = c2;
                }
                for (int i = 0; i < np; ++i) {
                    const int col = (i & 1) ? b : a;
                    M.B(OFF_AT + pl[i] * wmax + col) = 0xFF;
                    M.B(OFF_ATC + pc[i] * wmax + col) = 0xFF;
                }
                for (int i = 0; i < np; ++i) {
                    const int col = (i & 1) ? a : b;
                    M.B(OFF_AT + pl[i] * wmax + col) = (unsigned char)pc[i];
                    M.B(OFF_ATC + pc[i] * wmax + col) = (unsigned char)pl[i];
                }
                for (int i = 0; i < np; ++i) {
                    const int li = pl[i], ci = pc[i];
                    set_free(OFF_FL, li, a, M.B(OFF_AT + li * wmax + a) == 0xFF);
                    set_free(OFF_FL, li, b, M.B(OFF_AT + li * wmax + b) == 0xFF);
                    set_free(OFF_FC, ci, a, M.B(OFF_ATC + ci * wmax + a) == 0xFF);
                    set_free(OFF_FC, ci, b, M.B(OFF_ATC + ci * wmax + b) == 0xFF);
                }
                put(l, c, a);
            }
        }
    }
    // overflow entries: any free step of their lane (a conflict there is accepted)
    if (overflow) {
        for (int l = 0; l < NL; ++l)
            for (int c = 0; c < NL; ++c)
                for (int k = M.B(OFF_OVF + l * NL + c); k > 0; --k) {
                    const int t = first_free(OFF_FL, l);
                    if (t < 0) break;
                    M.B(OFF_AT + l * wmax + t) = (unsigned char)(c | 0x20);
                    set_free(OFF_FL, l, t, false);
                }
    }

    // ---- 3. codes ----
    for (int l = 0; l < NL; ++l) {
        for (int t = 0; t < W; ++t) {
            const int e = M.B(OFF_AT + l * wmax + t);
            unsigned char code;
            if (e == 0xFF) {
                int cf = -1;
                for (int c = 0; c < NL; ++c)
                    if ((M.Wd(OFF_FC + c * nw + (t >> 5)) >> (t & 31)) & 1u) { cf = c; break; }
                if (cf < 0) cf = l; else set_free(OFF_FC, cf, t, false);
                code = (unsigned char)(0x80 | cf);
            } else {
                const int c = e & MK, gb = (c + MK) & MK;
                const int ia = OFF_CNT + l * NL + c, ma = OFF_MOV + l * NL + c;
                const int ib = OFF_CNT + l * NL + gb, mb = OFF_MOV + l * NL + gb;
                if ((int)M.B(ia) - (int)M.B(ma) > 0) { M.B(ia) -= 1; code = (unsigned char)c; }
                else { M.B(mb) -= 1; M.B(ib) -= 1; code = (unsigned char)(gb | 0x10); }
            }
            out(l, t, code);
        }
    }
    return overflow;
}

template <class Mem, class Out>
SKM_HD static int skm_sched16(Mem &M, int W, int wmax, Out &out) { return skm_sched<16>(M, W, wmax, out); }

// ---------------------------------------------------------------------------------------------
// Second implementation of the same schedule: Birkhoff - von Neumann decomposition.
//
// skm_sched above keeps two NL x W tables per problem (2.5 KB at W = 78), which limits the device
// kernel to ~48 resident threads per SM.  Here the state is independent of W (4 NL^2 + 6 NL bytes):
// after the copies are balanced the (lane x class) count matrix is padded with free slots to a
// matrix whose rows and columns all sum to W.  Such a matrix is a sum of W permutation matrices
// (Birkhoff - von Neumann / Koenig): repeatedly take a perfect matching of the support (BFS
// augmenting paths on 16-bit adjacency masks; one exists while the matrix is regular), use it for
// mu = min multiplicity along it consecutive steps, subtract, and re-augment only the lanes whose
// cell ran empty.  Every step is a permutation lanes -> classes, i.e. conflict-free.
// Entries of a class that holds more than W entries (adversarial row patterns) cannot all get their
// own step; they are emitted in free slots of their lane (a bank conflict there, still correct).
// Same output codes as skm_sched.  Scratch: skm_bvn_bytes(NL) bytes, no words.
SKM_HD static inline int skm_bvn_bytes(int nl) { return 4 * nl * nl + 6 * nl; }

template <int NL, class Mem, class Out>
SKM_HD static int skm_sched_bvn(Mem &M, int W, Out &out)
{
    constexpr int MK = NL - 1, NONE = 0xFF;
    const int OFF_A = 0, OFF_B = NL * NL, OFF_P = 2 * NL * NL, OFF_O = 3 * NL * NL;
    const int OFF_ML = 4 * NL * NL, OFF_MC = OFF_ML + NL, OFF_PC = OFF_MC + NL, OFF_Q = OFF_PC + NL, OFF_ADJ = OFF_Q + NL;   // adj: 2 bytes per lane
    for (int i = NL * NL; i < 4 * NL * NL + 6 * NL; ++i) M.B(i) = 0;

    // ---- 1. balance the two copies (A = copy A of group c -> class c, B[l][c] = copy B of group c-1 -> class c) ----
    int T[NL], d[NL], x[NL];
    for (int c = 0; c < NL; ++c) {
        int s = 0;
        for (int l = 0; l < NL; ++l) s += M.B(OFF_A + l * NL + c);
        T[c] = s; d[c] = s; x[c] = 0;
    }
    for (int pass = 0; pass < 64; ++pass) {
        bool moved = false;
        for (int c = 0; c < NL; ++c) {
            const int ex = d[c] - W, avail = T[c] - x[c];
            if (ex > 0 && avail > 0) {
                const int mv = ex < avail ? ex : avail;
                x[c] += mv; d[c] -= mv; d[(c + 1) & MK] += mv;
                moved = true;
            }
        }
        if (!moved) break;
    }
    for (int c = 0; c < NL; ++c) {
        int rem = x[c];
        while (rem > 0) {
            bool any = false;
            for (int l = 0; l < NL && rem > 0; ++l) {
                const int a = M.B(OFF_A + l * NL + c);
                if (a > 0) {
                    M.B(OFF_A + l * NL + c) = (unsigned char)(a - 1);
                    M.B(OFF_B + l * NL + ((c + 1) & MK)) += 1;
                    --rem; any = true;
                }
            }
            if (!any) break;
        }
    }
    // ---- 2. classes that still hold more than W entries: the excess goes to free slots ----
    int overflow = 0;
    for (int c = 0; c < NL; ++c) {
        int e = d[c] - W;
        while (e > 0) {
            bool any = false;
            for (int l = 0; l < NL && e > 0; ++l) {
                const int a = M.B(OFF_A + l * NL + c), b = M.B(OFF_B + l * NL + c);
                if (a > 0) { M.B(OFF_A + l * NL + c) = (unsigned char)(a - 1); M.B(OFF_O + l * NL + c) += 1; }
                else if (b > 0) { M.B(OFF_B + l * NL + c) = (unsigned char)(b - 1); M.B(OFF_O + l * NL + ((c + MK) & MK)) += 1; }
                else continue;
                --e; ++overflow; any = true;
            }
            if (!any) break;
        }
        if (d[c] > W) d[c] = W;
    }
    // ---- 3. free slots: rows and columns are topped up to W (north-west corner rule) ----
    int rl[NL], ovl[NL];
    for (int l = 0; l < NL; ++l) {
        int s = 0, o = 0;
        for (int c = 0; c < NL; ++c) { s += M.B(OFF_A + l * NL + c) + M.B(OFF_B + l * NL + c); o += M.B(OFF_O + l * NL + c); }
        rl[l] = W - s; ovl[l] = o;
        if (rl[l] < 0) rl[l] = 0;                          // a lane longer than W: caller's bug, stay defined
    }
    {
        int i = 0, j = 0;
        int cc = W - d[0];
        while (i < NL && j < NL) {
            const int q = rl[i] < cc ? rl[i] : cc;
            if (q > 0) M.B(OFF_P + i * NL + j) += (unsigned char)q;
            rl[i] -= q; cc -= q;
            if (rl[i] == 0) ++i;
            else { ++j; if (j < NL) cc = W - d[j]; }
        }
    }
    // adjacency masks of the support
    auto cell = [&](int l, int c) -> int { return M.B(OFF_A + l * NL + c) + M.B(OFF_B + l * NL + c) + M.B(OFF_P + l * NL + c); };
    auto get_adj = [&](int l) -> uint32_t { return (uint32_t)M.B(OFF_ADJ + 2 * l) | ((uint32_t)M.B(OFF_ADJ + 2 * l + 1) << 8); };
    auto set_adj = [&](int l, uint32_t m) { M.B(OFF_ADJ + 2 * l) = (unsigned char)(m & 0xFF); M.B(OFF_ADJ + 2 * l + 1) = (unsigned char)(m >> 8); };
    for (int l = 0; l < NL; ++l) {
        uint32_t m = 0;
        for (int c = 0; c < NL; ++c) if (cell(l, c) > 0) m |= 1u << c;
        set_adj(l, m);
        M.B(OFF_ML + l) = NONE; M.B(OFF_MC + l) = NONE;
    }
    // BFS augmenting path from lane l0
    auto augment = [&](int l0) -> bool {
        uint32_t visited = 0;
        int qh = 0, qt = 0, found = -1;
        M.B(OFF_Q + qt++) = (unsigned char)l0;
        while (qh < qt && found < 0) {
            const int u = M.B(OFF_Q + qh++);
            uint32_t cand = get_adj(u) & ~visited;
            while (cand) {
                const int c = skm_ffs32(cand) - 1;
                cand &= cand - 1;
                visited |= 1u << c;
                M.B(OFF_PC + c) = (unsigned char)u;
                const int w = M.B(OFF_MC + c);
                if (w == NONE) { found = c; break; }
                M.B(OFF_Q + qt++) = (unsigned char)w;
            }
        }
        if (found < 0) return false;
        int c = found;
        for (;;) {
            const int u = M.B(OFF_PC + c);
            const int prev = M.B(OFF_ML + u);
            M.B(OFF_ML + u) = (unsigned char)c;
            M.B(OFF_MC + c) = (unsigned char)u;
            if (u == l0) break;
            c = prev;
        }
        return true;
    };
    for (int l = 0; l < NL; ++l) augment(l);

    // ---- 4. decomposition ----
    int t = 0;
    while (t < W) {
        int mu = W - t;
        for (int l = 0; l < NL; ++l) {
            const int c = M.B(OFF_ML + l);
            const int v = c == NONE ? 1 : cell(l, c);
            if (v < mu) mu = v;
        }
        if (mu < 1) mu = 1;
        for (int l = 0; l < NL; ++l) {
            int c = M.B(OFF_ML + l);
            for (int s = 0; s < mu; ++s) {
                unsigned char code;
                if (c == NONE) {
                    // irregular remainder (cannot happen for a padded regular matrix): flush any real entry
                    int g = -1, which = 0;
                    for (int k = 0; k < NL && g < 0; ++k) {
                        if (M.B(OFF_A + l * NL + k) > 0) { g = k; which = 0; }
                        else if (M.B(OFF_B + l * NL + k) > 0) { g = k; which = 1; }
                        else if (M.B(OFF_O + l * NL + k) > 0) { g = k; which = 2; }
                    }
                    if (g < 0) code = (unsigned char)(0x80 | l);
                    else if (which == 0) { M.B(OFF_A + l * NL + g) -= 1; code = (unsigned char)g; ++overflow; }
                    else if (which == 1) { M.B(OFF_B + l * NL + g) -= 1; code = (unsigned char)(((g + MK) & MK) | 0x10); ++overflow; }
                    else { M.B(OFF_O + l * NL + g) -= 1; ovl[l] -= 1; code = (unsigned char)g; }
                } else if (M.B(OFF_A + l * NL + c) > 0) { M.B(OFF_A + l * NL + c) -= 1; code = (unsigned char)c; }
                else if (M.B(OFF_B + l * NL + c) > 0) { M.B(OFF_B + l * NL + c) -= 1; code = (unsigned char)(((c + MK) & MK) | 0x10); }
                else {
                    M.B(OFF_P + l * NL + c) -= 1;
                    if (ovl[l] > 0) {                       // a free slot carries one of the lane's overflow entries
                        int g = 0;
                        while (g < NL - 1 && M.B(OFF_O + l * NL + g) == 0) ++g;
                        M.B(OFF_O + l * NL + g) -= 1; ovl[l] -= 1;
                        code = (unsigned char)g;
                    } else code = (unsigned char)(0x80 | c);
                }
                out(l, t + s, code);
            }
        }
        t += mu;
        if (t >= W) break;
        // lanes whose cell ran empty lose their match and are re-augmented
        uint32_t redo = 0;
        for (int l = 0; l < NL; ++l) {
            const int c = M.B(OFF_ML + l);
            if (c == NONE) { redo |= 1u << l; continue; }
            if (cell(l, c) == 0) {
                set_adj(l, get_adj(l) & ~(1u << c));
                M.B(OFF_MC + c) = NONE; M.B(OFF_ML + l) = NONE;
                redo |= 1u << l;
            }
        }
        while (redo) {
            const int l = skm_ffs32(redo) - 1;
            redo &= redo - 1;
            augment(l);
        }
    }
    return overflow;
}
