// spmm_kernels.cuh -- device code of the forward SpMM (templates) and the per-reduction kernel
// pickers.  Included by spmm_inst_{sum,max,min}.cu, which instantiate one reduction each so
// the three translation units compile in parallel; spmm_fwd.cu holds the host-side dispatch.
#pragma once
#ifndef ISPLIB_LEAN_PREFETCH
#define ISPLIB_LEAN_PREFETCH 1
#endif
#include "common.cuh"
#include <float.h>
#include <limits.h>

namespace isplib {

typedef void (*SegKernel)(const SpmmParams);

// ------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------
template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

#ifndef ISPLIB_X_EVICT_LAST
#define ISPLIB_X_EVICT_LAST 0
#endif

#if ISPLIB_X_EVICT_LAST
// L2 cache policy for the gathered X rows: keep them (evict_last) while the index stream and
// the output go through with evict-first hints, so a K slab of X survives in L2.
__device__ __forceinline__ unsigned long long x_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#define ISPLIB_XPOL_DECL const unsigned long long xpol = x_policy();
#define ISPLIB_XPOL_ARG , xpol
#define ISPLIB_XPOL_PARAM , unsigned long long xpol
#else
#define ISPLIB_XPOL_DECL
#define ISPLIB_XPOL_ARG
#define ISPLIB_XPOL_PARAM
#endif

template <int VEC>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[VEC] ISPLIB_XPOL_PARAM) {
#if ISPLIB_X_EVICT_LAST
    if constexpr (VEC == 4) {
        asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(xpol));
    } else {
        asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[0]) : "l"(p), "l"(xpol));
    }
#else
    if constexpr (VEC == 8) {
        // 256-bit load (sm_100: LDG.E.256): measured +25 % random-row-gather bandwidth over
        // 128-bit loads out of L2 (profiles/r1_l2probe.txt, "gather256")
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p));
    } else if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
#endif
}

template <int VEC>
__device__ __forceinline__ void store_vec_f(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 8) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
    } else if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else {
        __stcs(p, v[0]);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec_i64(long long* p, const long long (&v)[VEC]) {
    if constexpr (VEC == 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q) __stcs(reinterpret_cast<longlong2*>(p) + q, make_longlong2(v[2 * q], v[2 * q + 1]));
    } else if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<longlong2*>(p), make_longlong2(v[0], v[1]));
        __stcs(reinterpret_cast<longlong2*>(p) + 1, make_longlong2(v[2], v[3]));
    } else {
        __stcs(p, v[0]);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec_i32(int* p, const int (&v)[VEC]) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]);
    } else {
        *p = v[0];
    }
}

template <int OP> __device__ __forceinline__ float init_value() {
    // csrc/fusedmm.cpp:147-152: zeros / numeric_limits<float>::lowest() / ::max()
    return OP == OP_SUM ? 0.f : (OP == OP_MAX ? -FLT_MAX : FLT_MAX);
}

constexpr int kNoArg = INT_MAX;  // "no entry won yet" (local edge ids are < 2^31-1)

// strict compare, first (smallest edge id) wins: mirrors oracle/fusedmm_oracle.c
template <int OP>
__device__ __forceinline__ bool better(float cand, float cur) {
    return OP == OP_MAX ? (cand > cur) : (cand < cur);
}
template <int OP, typename IdT>
__device__ __forceinline__ bool better_lex(float cand, IdT cand_id, float cur, IdT cur_id) {
    return better<OP>(cand, cur) || (cand == cur && cand_id < cur_id);
}

// ------------------------------------------------------------------------------------
// finalisation of one vector of one output row: single-segment rows directly, split rows by
// the last-arriving segment warp after the in-order merge
// ------------------------------------------------------------------------------------
// fused caller epilogue, applied to the finished row vector:
//   out = relu( out + addend_scale * addend[row, :] + bias[:] )
// GCN: + bias, ReLU (tests/cpu/gcn-sparse.py:61-68); GIN: (1 + eps) * x_i + sum_j x_j
// (gin-sparse.py:73-78) with addend = x; every part optional.
template <int VEC>
__device__ __forceinline__ void apply_epilogue(const SpmmParams& p, int row, int kk, int nvalid, float (&acc)[VEC]) {
    if (p.addend) {
        const float* a = p.addend + (size_t)row * (size_t)p.ld_addend + kk;
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (v < nvalid) acc[v] = fmaf(p.addend_scale, __ldg(a + v), acc[v]);
    }
    if (p.bias) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (v < nvalid) acc[v] += __ldg(p.bias + kk + v);
    }
    if (p.flags & ISPLIB_FLAG_RELU) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaxf(acc[v], 0.f);
    }
}

template <int OP, int VEC>
__device__ __forceinline__ void finalize_store(const SpmmParams& p, int row, int deg, int kk,
                                               float (&acc)[VEC], int (&arg)[VEC]) {
    const size_t o = (size_t)row * (size_t)p.ldo + (size_t)kk;
    // 16-byte stores need an aligned out row and the whole vector inside [0, k); otherwise
    // (K % 4 != 0: the last vector of a row, or an unaligned out) fall back to scalars
    const bool vec_ok = (VEC == 1) || (p.vec_store && kk + VEC <= p.k);
    const int nvalid = min(VEC, p.k - kk);
    if constexpr (OP == OP_SUM) {
        if (p.flags & ISPLIB_FLAG_ACCUMULATE) {
            float prev[VEC];
            if (VEC >= 4 && vec_ok) {
#pragma unroll
                for (int q = 0; q < VEC / 4; ++q) {
                    const float4 t = *(reinterpret_cast<const float4*>(p.out + o) + q);
                    prev[(4 * q) % VEC] = t.x; prev[(4 * q + 1) % VEC] = t.y; prev[(4 * q + 2) % VEC] = t.z; prev[(4 * q + 3) % VEC] = t.w;
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) prev[v] = (v < nvalid) ? p.out[o + v] : 0.f;
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = prev[v] + acc[v];
        }
        if (p.div_mode) {
            const float d = (p.div_mode == 2) ? __ldg(p.row_div + row) : (float)max(deg, 1);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = __fdiv_rn(acc[v], d);
        }
        if (p.has_epilogue) apply_epilogue<VEC>(p, row, kk, nvalid, acc);
        if (vec_ok) {
            store_vec_f<VEC>(p.out + o, acc);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) if (v < nvalid) __stcs(p.out + o + v, acc[v]);
        }
    } else {
        long long gid[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (arg[v] == kNoArg) gid[v] = p.arg_sentinel;
            else gid[v] = p.edge_ids ? (long long)__ldg(p.edge_ids + arg[v]) : (long long)arg[v];
        }
        if (p.flags & ISPLIB_FLAG_ACCUMULATE) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (v < nvalid) {
                    const float pv = p.out[o + v];
                    const long long pa = p.arg_out[o + v];
                    // a previous block that found no entry (sentinel) holds no candidate: ours is
                    // taken as is (its stored value may be the EMPTY_ZERO placeholder, not a product);
                    // otherwise the previous block wins unless ours is strictly better / equal with
                    // a smaller edge id
                    if (pa != p.arg_sentinel && !better_lex<OP, long long>(acc[v], gid[v], pv, pa)) { acc[v] = pv; gid[v] = pa; }
                }
            }
        }
        if (p.arg_col) {
            // the winner's column (and value) for the backward scatter, so that it reads two
            // coalesced 4-byte streams instead of gathering col[arg] / val[arg] by 32-byte sectors
            // (csrc/fusedmm.cpp:432-441 does the same gathers with index_select)
            int cw[VEC];
            float aw[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const bool has = (arg[v] != kNoArg);
                cw[v] = has ? __ldg(p.col + arg[v]) : -1;
                aw[v] = (has && p.val) ? __ldg(p.val + arg[v]) : 1.f;
            }
            if (VEC >= 4 && vec_ok) {
#pragma unroll
                for (int q = 0; q < VEC / 4; ++q) {
                    __stcs(reinterpret_cast<int4*>(p.arg_col + o) + q, make_int4(cw[(4 * q) % VEC], cw[(4 * q + 1) % VEC], cw[(4 * q + 2) % VEC], cw[(4 * q + 3) % VEC]));
                    if (p.arg_val)
                        __stcs(reinterpret_cast<float4*>(p.arg_val + o) + q, make_float4(aw[(4 * q) % VEC], aw[(4 * q + 1) % VEC], aw[(4 * q + 2) % VEC], aw[(4 * q + 3) % VEC]));
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (v < nvalid) { __stcs(p.arg_col + o + v, cw[v]); if (p.arg_val) __stcs(p.arg_val + o + v, aw[v]); }
            }
        }
        if (p.flags & ISPLIB_FLAG_EMPTY_ZERO) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) if (gid[v] == p.arg_sentinel) acc[v] = 0.f;
        }
        if (p.has_epilogue) apply_epilogue<VEC>(p, row, kk, nvalid, acc);
        if (vec_ok) {
            store_vec_f<VEC>(p.out + o, acc);
            store_vec_i64<VEC>(p.arg_out + o, gid);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (v < nvalid) { __stcs(p.out + o + v, acc[v]); __stcs(p.arg_out + o + v, gid[v]); }
        }
    }
}

// ------------------------------------------------------------------------------------
// what every gather strategy ends with: merge the lane groups, then finalise the row (single
// segment) or publish a partial and let the last-arriving segment warp merge the row
// ------------------------------------------------------------------------------------
template <int OP, int VEC, int G, int LPL>
__device__ __forceinline__ void finish_item(const SpmmParams& p, const int lane, const int row, const int eb,
                                            const int ee, const int slot_id, const int (&koff)[LPL],
                                            const bool (&kok)[LPL], float (&acc)[LPL][VEC], int (&arg)[LPL][VEC]) {
    constexpr int NG = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    const int g = lane / G;
    // merge the NG lane groups (they hold interleaved entries of the same segment)
    if constexpr (NG > 1) {
#pragma unroll
        for (int off = G; off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < LPL; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float ov = __shfl_xor_sync(FULL, acc[j][v], off);
                    if constexpr (OP == OP_SUM) {
                        acc[j][v] += ov;
                    } else {
                        const int oa = __shfl_xor_sync(FULL, arg[j][v], off);
                        if (better_lex<OP, int>(ov, oa, acc[j][v], arg[j][v])) { acc[j][v] = ov; arg[j][v] = oa; }
                    }
                }
        }
    }
    const bool writer = (NG == 1) || (g == 0);   // group 0 holds the merged result

    if (slot_id < 0) {   // the row fits this one segment: its degree is ee - eb
        if (writer) {
#pragma unroll
            for (int j = 0; j < LPL; ++j)
                if (kok[j]) finalize_store<OP, VEC>(p, row, ee - eb, koff[j], acc[j], arg[j]);
        }
        return;
    }

    // ---- split row: publish this segment's partial; the LAST segment warp to arrive merges
    // all of the row's partials in segment order (the threadFenceReduction pattern: nobody
    // waits, the order of the merge is fixed, so the result is deterministic and max/min/arg
    // stay bit-exact) and finalises the row.  No second kernel launch.
    const int pbase = __ldg(p.part_off + row);
    const int nseg = __ldg(p.seg_off + row + 1) - __ldg(p.seg_off + row);
    const int rb = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
    if (writer) {
        const size_t slot = (size_t)slot_id;
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
            if (kok[j]) {
                const size_t o = slot * (size_t)p.kp + (size_t)koff[j];
                if constexpr (VEC >= 4) {
#pragma unroll
                    for (int q = 0; q < VEC / 4; ++q) {
                        __stcg(reinterpret_cast<float4*>(p.part_val + o) + q,
                               make_float4(acc[j][(4 * q) % VEC], acc[j][(4 * q + 1) % VEC], acc[j][(4 * q + 2) % VEC], acc[j][(4 * q + 3) % VEC]));
                        if constexpr (OP != OP_SUM)
                            __stcg(reinterpret_cast<int4*>(p.part_arg + o) + q,
                                   make_int4(arg[j][(4 * q) % VEC], arg[j][(4 * q + 1) % VEC], arg[j][(4 * q + 2) % VEC], arg[j][(4 * q + 3) % VEC]));
                    }
                } else {
                    __stcg(p.part_val + o, acc[j][0]);
                    if constexpr (OP != OP_SUM) __stcg(p.part_arg + o, arg[j][0]);
                }
            }
        }
    }
    __threadfence();
    __syncwarp();
    int* const ticket_ptr = p.row_ticket + (size_t)(blockIdx.y + p.tile_base) * (size_t)p.ticket_stride + (pbase >> 1);
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(ticket_ptr, 1);
    ticket = __shfl_sync(FULL, ticket, 0);
    if (ticket != nseg - 1) return;
    __threadfence();
    if (lane == 0) *ticket_ptr = 0;   // leave the counters zeroed for the next launch
    if (!writer) return;
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        if (!kok[j]) continue;
        float macc[VEC];
        int marg[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { macc[v] = init_value<OP>(); marg[v] = kNoArg; }
        const size_t o0 = (size_t)pbase * (size_t)p.kp + (size_t)koff[j];
#pragma unroll 4
        for (int t = 0; t < nseg; ++t) {
            const size_t o = o0 + (size_t)t * (size_t)p.kp;
            float pv[VEC];
            int pa[VEC];
            if constexpr (VEC >= 4) {
#pragma unroll
                for (int h = 0; h < VEC / 4; ++h) {
                    const float4 q = __ldcg(reinterpret_cast<const float4*>(p.part_val + o) + h);
                    pv[(4 * h) % VEC] = q.x; pv[(4 * h + 1) % VEC] = q.y; pv[(4 * h + 2) % VEC] = q.z; pv[(4 * h + 3) % VEC] = q.w;
                    if constexpr (OP != OP_SUM) {
                        const int4 r = __ldcg(reinterpret_cast<const int4*>(p.part_arg + o) + h);
                        pa[(4 * h) % VEC] = r.x; pa[(4 * h + 1) % VEC] = r.y; pa[(4 * h + 2) % VEC] = r.z; pa[(4 * h + 3) % VEC] = r.w;
                    }
                }
            } else {
                pv[0] = __ldcg(p.part_val + o);
                if constexpr (OP != OP_SUM) pa[0] = __ldcg(p.part_arg + o);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if constexpr (OP == OP_SUM) {
                    macc[v] += pv[v];
                } else {
                    // segments are in increasing edge order: strict compare keeps the first
                    if (better<OP>(pv[v], macc[v])) { macc[v] = pv[v]; marg[v] = pa[v]; }
                }
            }
        }
        finalize_store<OP, VEC>(p, row, re - rb, koff[j], macc, marg);
    }
}

// ------------------------------------------------------------------------------------
// main kernel: one warp = one row segment x one K tile
// ------------------------------------------------------------------------------------
// PARTIAL: some lanes of the widest K tile fall outside [k0, kend) (K = 100, 200, 47 ...).
// Those lanes gather the tile's first vector instead (same 16 bytes a valid lane reads, so
// no extra traffic, and no predicate or zero-fill on the hot loads) and never store.
template <int OP, int VEC, int G, int LPL, int U, bool PARTIAL>
__global__ void __launch_bounds__(256)
spmm_seg_kernel(const __grid_constant__ SpmmParams p) {
    constexpr int NG = 32 / G;         // lane groups per warp = entries gathered per step
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(32 % (NG * U) == 0, "a 32-entry chunk must be a whole number of steps");

    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= p.num_items) return;

    // one 16-byte descriptor per work item {row, first entry, end entry, partial slot or -1}:
    // a single load instead of the item_row -> seg_off -> rowptr chain, which matters for
    // short rows (products-shape: ~50 entries per row, the chain was ~1/3 of a warp's life)
    const int4 desc = __ldg(p.item_desc + item);
    const int row = desc.x, eb = desc.y, ee = desc.z, slot_id = desc.w;

    const int g = lane / G;
    const int lg = lane % G;
    const int k0 = (blockIdx.y + p.tile_base) * p.tile_w;
    // VEC=4 loads may read up to 3 padding floats past K (the launcher checked ldx >= roundup4(K))
    const int keff = (VEC > 1) ? ((p.k + VEC - 1) & ~(VEC - 1)) : p.k;
    const int kend = min(keff, k0 + p.tile_w);

    int koff[LPL];
    bool kok[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
        koff[j] = k0 + (lg + j * G) * VEC;
        kok[j] = koff[j] < kend;
    }

    // gather address = lane base + col * ldx_bytes: one IMAD.WIDE.U32 per gathered row
    // instead of a 64x64-bit multiply; further vectors of the lane sit at immediate offsets
    const char* xlane[PARTIAL ? LPL : 1];
    xlane[0] = reinterpret_cast<const char*>(p.x + (kok[0] ? koff[0] : k0));
    if constexpr (PARTIAL) {
#pragma unroll
        for (int j = 1; j < LPL; ++j) xlane[j] = reinterpret_cast<const char*>(p.x + (kok[j] ? koff[j] : k0));
    }
    const unsigned ldxb = (unsigned)p.ldx * 4u;

    float acc[LPL][VEC];
    int arg[LPL][VEC];
#pragma unroll
    for (int j = 0; j < LPL; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { acc[j][v] = init_value<OP>(); arg[j][v] = kNoArg; }

    const bool has_val = (p.val != nullptr);
    ISPLIB_XPOL_DECL

    // one 32-entry chunk of (col, val) per lane, one chunk prefetched ahead
    unsigned c_next = 0;
    float a_next = 0.f;
    if (eb + lane < ee) {
        c_next = (unsigned)__ldcs(p.col + eb + lane);
        a_next = has_val ? __ldcs(p.val + eb + lane) : 1.f;
    }

    for (int e0 = eb; e0 < ee; e0 += 32) {
        const unsigned c = c_next;
        const float a = a_next;
        const int cnt = min(32, ee - e0);
        {
            const int en = e0 + 32 + lane;
            if (en < ee) {
                c_next = (unsigned)__ldcs(p.col + en);
                a_next = has_val ? __ldcs(p.val + en) : 1.f;
            }
        }
        if (cnt == 32) {
#pragma unroll
            for (int t = 0; t < 32; t += NG * U) {
                float xv[U][LPL][VEC];
                float aa[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = t + u * NG + g;
                    const unsigned cc = __shfl_sync(FULL, c, idx);
                    aa[u] = __shfl_sync(FULL, a, idx);
                    const unsigned long long off = (unsigned long long)cc * ldxb;
#pragma unroll
                    for (int j = 0; j < LPL; ++j) {
                        if constexpr (PARTIAL) load_vec<VEC>(reinterpret_cast<const float*>(xlane[j] + off), xv[u][j] ISPLIB_XPOL_ARG);
                        else load_vec<VEC>(reinterpret_cast<const float*>(xlane[0] + off) + j * G * VEC, xv[u][j] ISPLIB_XPOL_ARG);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = e0 + t + u * NG + g;
#pragma unroll
                    for (int j = 0; j < LPL; ++j)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            if constexpr (OP == OP_SUM) {
                                acc[j][v] = fmaf(aa[u], xv[u][j][v], acc[j][v]);
                            } else {
                                const float tt = __fmul_rn(aa[u], xv[u][j][v]);
                                if (better<OP>(tt, acc[j][v])) { acc[j][v] = tt; arg[j][v] = e; }
                            }
                        }
                }
            }
        } else {
            // ragged tail of the segment: same steps, predicated per entry
            for (int t = 0; t < cnt; t += NG * U) {
                float xv[U][LPL][VEC];
                float aa[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = t + u * NG + g;
                    ok[u] = idx < cnt;
                    const unsigned cc = __shfl_sync(FULL, c, idx & 31);
                    aa[u] = __shfl_sync(FULL, a, idx & 31);
                    const unsigned long long off = (unsigned long long)cc * ldxb;
                    if (ok[u]) {
#pragma unroll
                        for (int j = 0; j < LPL; ++j) {
                            if constexpr (PARTIAL) load_vec<VEC>(reinterpret_cast<const float*>(xlane[j] + off), xv[u][j] ISPLIB_XPOL_ARG);
                            else load_vec<VEC>(reinterpret_cast<const float*>(xlane[0] + off) + j * G * VEC, xv[u][j] ISPLIB_XPOL_ARG);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = e0 + t + u * NG + g;
                    if (ok[u]) {
#pragma unroll
                        for (int j = 0; j < LPL; ++j)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                if constexpr (OP == OP_SUM) {
                                    acc[j][v] = fmaf(aa[u], xv[u][j][v], acc[j][v]);
                                } else {
                                    const float tt = __fmul_rn(aa[u], xv[u][j][v]);
                                    if (better<OP>(tt, acc[j][v])) { acc[j][v] = tt; arg[j][v] = e; }
                                }
                            }
                    }
                }
            }
        }
    }

    finish_item<OP, VEC, G, LPL>(p, lane, row, eb, ee, slot_id, koff, kok, acc, arg);
}

// ------------------------------------------------------------------------------------
// bulk-copy gather variant: the dense rows are fetched by the TMA engine
// (cp.async.bulk global -> shared, completion on an mbarrier) instead of by LDG into registers.
// Every lane issues the bulk copy of ONE row (no address math or data registers per 16 bytes),
// SE rows per stage and STAGES stages per warp are in flight -- bytes in flight are bounded by
// shared memory (24-48 KB per warp), not by registers -- and the FMA / compare loop reads the
// rows back with conflict-free LDS.128.  Warp-private ring: no __syncthreads, each warp owns
// its stage buffers and its mbarriers.  Same work items, same finish_item as the LDG kernel.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kBulkSE = 16;   // rows (entries) per stage

template <int OP, int G, int LPL, int STAGES>
__global__ void __launch_bounds__(256)
spmm_bulk_kernel(const __grid_constant__ SpmmParams p) {
    constexpr int VEC = 4;
    constexpr int NG = 32 / G;
    constexpr int SE = kBulkSE;
    constexpr int ROW_STRIDE = G * LPL * 16;          // bytes of one staged row slot
    constexpr int STAGE_BYTES = SE * ROW_STRIDE;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int item = blockIdx.x * nwarps + wib;
    if (item >= p.num_items) return;

    unsigned char* my = smem_raw + (size_t)wib * (STAGES * STAGE_BYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)nwarps * (STAGES * STAGE_BYTES)) + wib * STAGES;
    const unsigned buf0 = smem_u32(my);
    const unsigned bar0 = smem_u32(bars);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(bar0 + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int4 desc = __ldg(p.item_desc + item);
    const int row = desc.x, eb = desc.y, ee = desc.z, slot_id = desc.w;
    const int g = lane / G, lg = lane % G;
    const int k0 = (blockIdx.y + p.tile_base) * p.tile_w;
    const int keff = (p.k + 3) & ~3;
    const int kend = min(keff, k0 + p.tile_w);
    const unsigned row_bytes = (unsigned)(kend - k0) * 4u;

    int koff[LPL];
    bool kok[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) { koff[j] = k0 + (lg + j * G) * VEC; kok[j] = koff[j] < kend; }
    float acc[LPL][VEC];
    int arg[LPL][VEC];
#pragma unroll
    for (int j = 0; j < LPL; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { acc[j][v] = init_value<OP>(); arg[j][v] = kNoArg; }

    const bool has_val = (p.val != nullptr);
    const char* const xbase = reinterpret_cast<const char*>(p.x + k0);
    const unsigned ldxb = (unsigned)p.ldx * 4u;
    const int nstages = (ee - eb + SE - 1) / SE;
    float a_reg[STAGES];

    auto issue = [&](int s, int i) {
        if (s >= nstages) return;
        const int e0 = eb + s * SE;
        const int cnt = min(SE, ee - e0);
        unsigned c = 0;
        a_reg[i] = 0.f;
        if (lane < cnt) {
            c = (unsigned)__ldcs(p.col + e0 + lane);
            a_reg[i] = has_val ? __ldcs(p.val + e0 + lane) : 1.f;
        }
        if (lane == 0) mbar_arrive_expect_tx(bar0 + 8u * i, (unsigned)cnt * row_bytes);
        __syncwarp();
        if (lane < cnt)
            bulk_g2s(buf0 + (unsigned)(i * STAGE_BYTES + lane * ROW_STRIDE), xbase + (unsigned long long)c * ldxb,
                     row_bytes, bar0 + 8u * i);
    };

    // prologue: fill the ring
#pragma unroll
    for (int i = 0; i < STAGES; ++i) issue(i, i);

    for (int base = 0; base < nstages; base += STAGES) {
        const unsigned parity = (unsigned)((base / STAGES) & 1);
#pragma unroll
        for (int i = 0; i < STAGES; ++i) {
            const int s = base + i;
            if (s < nstages) {                      // warp-uniform
                const int e0 = eb + s * SE;
                const int cnt = min(SE, ee - e0);
                while (!mbar_try_wait(bar0 + 8u * i, parity)) { }
                const unsigned char* stage = my + i * STAGE_BYTES;
#pragma unroll
                for (int t = 0; t < SE; t += NG) {
                    const int idx = t + g;
                    const float a = __shfl_sync(FULL, a_reg[i], idx);
                    if (idx < cnt) {
                        const int e = e0 + idx;
#pragma unroll
                        for (int j = 0; j < LPL; ++j) {
                            if (kok[j]) {
                                const float4 q = *reinterpret_cast<const float4*>(stage + idx * ROW_STRIDE + (lg + j * G) * 16);
                                const float xv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    if constexpr (OP == OP_SUM) {
                                        acc[j][v] = fmaf(a, xv[v], acc[j][v]);
                                    } else {
                                        const float tt = __fmul_rn(a, xv[v]);
                                        if (better<OP>(tt, acc[j][v])) { acc[j][v] = tt; arg[j][v] = e; }
                                    }
                                }
                            }
                        }
                    }
                }
                // every lane is done reading this stage before the TMA engine may overwrite it
                __syncwarp();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(s + STAGES, i);
            }
        }
    }
    finish_item<OP, VEC, G, LPL>(p, lane, row, eb, ee, slot_id, koff, kok, acc, arg);
}

// ------------------------------------------------------------------------------------
// lean kernel (method 5: VEC = 8, 32-byte gathers `lean256/*`; method 6: VEC = 4 `lean128/*`):
// the 32-byte gather pays off only if four of them stay in flight per lane at >= 32 warps/SM,
// i.e. within 64 registers.  This body drops what the general kernel carries (partial-tile
// bookkeeping, index prefetch, staged edge values) and keeps the step loop rolled so only
// U x VEC staging registers are live.
// Whole tiles only (K % tile == 0); a tile narrower than G x VEC floats (RAGGED, e.g. K = 200
// -> 25 of 32 lanes) parks the surplus lanes on the tile's first vector (same sectors as lane
// 0, so no extra traffic) and keeps them out of the stores.  max/min carry VEC more registers
// (arg): 80 with VEC = 8 (24 warps/SM), 56 with VEC = 4 (36 warps/SM).
// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// fused all-gather (GatherParams, common.cuh): the push role and the consumer-side wait
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_sys_add(unsigned* p, unsigned v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr unsigned long long kGatherTimeoutNs = 4000000000ull;   // 4 s: a peer that never shows up must not hang the GPU

// spin until *word has reached `target` (wrap-safe); false on timeout
__device__ __forceinline__ bool wait_reached(const unsigned* word, unsigned target, unsigned* status) {
    const unsigned long long t0 = global_timer_ns();
    unsigned ns = 32;
    while (true) {
        const unsigned v = ld_acquire_sys(word);
        if ((int)(v - target) >= 0) return true;
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        if (global_timer_ns() - t0 > kGatherTimeoutNs) { atomicExch(status, 1u); return false; }
    }
}

// The first copy_ctas CTAs: push my slice to every peer.  Each vector is loaded ONCE (local, L2) and
// stored to all peers (posted NVLink writes); 4 vectors per thread in flight.
static __device__ __noinline__ void gather_push_role(const GatherParams& G) {
    const int cta = blockIdx.x, nc = G.copy_ctas, tid = threadIdx.x, nt = blockDim.x;
    // credit: "I have started step `epoch`" -- the step before has released my buffer of the OTHER
    // parity, so peers may push step epoch + 1 into it
    if (cta == 0 && tid < G.n_dst) st_release_sys(G.peer_credit[tid] + G.my_rank, G.epoch);
    if (G.phase == 2) return;
    // nobody may overwrite a buffer its owner still reads: wait for every peer's credit (steady state: long there)
    if (tid < G.n_dst) wait_reached(G.credit_local + G.dst_rank[tid], G.epoch - 1u, G.status);
    __syncthreads();
    const float4* __restrict__ own = reinterpret_cast<const float4*>(G.own);
    const int nd = G.n_dst;
    if (G.tile_vec4 > 0) {
        // ---- tile mode: K tile after K tile (a [slice_rows x tile] block, row pitch row_vec4), to every peer
        const long long n = G.slice_rows * (long long)G.tile_vec4;
        const long long chunk = ((n + nc - 1) / nc + 3) & ~3ll;
        const long long b = min(n, (long long)cta * chunk), e = min(n, b + chunk);
        const int tv = G.tile_vec4, rv = G.row_vec4;
        for (int t = 0; t < G.n_groups; ++t) {
            for (long long i = b + tid; i < e; i += (long long)nt * 4) {
                float4 v[4];
                long long off[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long j = i + (long long)u * nt;
                    const long long r = j / tv;
                    off[u] = r * rv + (j - r * tv) + (long long)t * tv;
                    if (j < e) v[u] = __ldg(own + off[u]);
                }
                for (int s = 0; s < nd; ++s) {
                    float4* __restrict__ dst = reinterpret_cast<float4*>(G.dst[s]);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i + (long long)u * nt < e) dst[off[u]] = v[u];
                }
            }
            __threadfence_system();          // my stores have reached the peers ...
            __syncthreads();                 // ... for every thread of the CTA ...
            if (tid < nd) red_release_sys_add(G.peer_arrive[tid] + t, 1u);   // ... before the arrival is published there
        }
        return;
    }
    // ---- owner mode: the whole slice, peer after peer in the order of the group it belongs to THERE
    const long long n = G.slice_vec4;
    const long long chunk = ((n + nc - 1) / nc + 3) & ~3ll;
    const long long b = min(n, (long long)cta * chunk), e = min(n, b + chunk);
    for (int s = 0; s < nd; ++s) {
        float4* __restrict__ dst = reinterpret_cast<float4*>(G.dst[s]);
        for (long long i = b + tid; i < e; i += (long long)nt * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i + (long long)u * nt < e) v[u] = __ldg(own + i + (long long)u * nt);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i + (long long)u * nt < e) dst[i + (long long)u * nt] = v[u];
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) red_release_sys_add(G.peer_arrive[s] + G.dst_group[s], 1u);
    }
}

// consumer side.  Tile mode: every item of a CTA gathers remote rows of the same K tile, so ONE
// thread per CTA polls the tile's arrival counter (4x fewer pollers than one per warp, and a backoff
// up to 4 us: thousands of warps hammering one L2 sector delay the very RED they are waiting for).
__device__ __forceinline__ void gather_wait_for_tile(const GatherParams& G, int tile) {
    if (threadIdx.x == 0) {
        const unsigned long long t0 = global_timer_ns();
        unsigned ns = 64;
        while ((int)(ld_acquire_sys(G.arrive_local + tile) - G.arrive_target[tile]) < 0) {
            __nanosleep(ns);
            if (ns < 4096) ns <<= 1;
            if (global_timer_ns() - t0 > kGatherTimeoutNs) { atomicExch(G.status, 1u); break; }
        }
    }
    __syncthreads();
}
// Owner mode: the arrival group of a work item (group 0 = the rank's own slice: nothing to wait for)
__device__ __forceinline__ void gather_wait_for_item(const GatherParams& G, int item, int lane) {
    int g = 0;
    while (g < G.n_groups - 1 && item >= G.group_item_end[g]) ++g;
    if (g == 0) return;
    if (lane == 0) wait_reached(G.arrive_local + g, G.arrive_target[g], G.status);
    __syncwarp();
}

// max / min update of one accumulator: if (a * x is strictly better than acc) { acc = a * x; arg = e; }
// The compare and the arg select issue on the ALU pipe; writing acc as a PREDICATED MULTIPLY (the same
// product again, bit-identical) instead of a select puts that write on the FMA pipe next to the first
// multiply: 2 + 2 instructions per element on the two half-rate pipes instead of 3 + 1 (both pipes
// issue one warp instruction every 2 cycles per SM sub-partition; the 3 + 1 split made the ALU pipe
// the limiter of the max/min kernels).
#ifndef ISPLIB_ARGTRACK_PREDMUL
#define ISPLIB_ARGTRACK_PREDMUL 1
#endif
template <int OP>
__device__ __forceinline__ void track_better(float a, float x, int e, float nz, float& acc, int& arg) {
#if ISPLIB_ARGTRACK_PREDMUL
    // nz == -0.0f, but opaque to the compiler (derived from a launch parameter): fma(a, x, -0.0) is
    // a * x bit for bit (one rounding, the sign of a zero product is kept), yet ptxas cannot merge it
    // with the multiply above it into one FMUL + FSEL
    if constexpr (OP == OP_MAX) {
        asm("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\t"
            "mul.rn.f32 t, %2, %3;\n\t"
            "setp.gt.f32 p, t, %0;\n\t"
            "@p fma.rn.f32 %0, %2, %3, %5;\n\t"
            "selp.s32 %1, %4, %1, p;\n\t}"
            : "+f"(acc), "+r"(arg) : "f"(a), "f"(x), "r"(e), "f"(nz));
    } else {
        asm("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\t"
            "mul.rn.f32 t, %2, %3;\n\t"
            "setp.lt.f32 p, t, %0;\n\t"
            "@p fma.rn.f32 %0, %2, %3, %5;\n\t"
            "selp.s32 %1, %4, %1, p;\n\t}"
            : "+f"(acc), "+r"(arg) : "f"(a), "f"(x), "r"(e), "f"(nz));
    }
#else
    (void)nz;
    const float tt = __fmul_rn(a, x);
    if (better<OP>(tt, acc)) { acc = tt; arg = e; }
#endif
}

// NOVAL (max / min only): val == NULL (SAGE / GIN drop the values) -- no multiply and no value
// shuffle at all in the hot loop.
//
// Tried in round 2 and removed again (profiles/r2_kbench_max_stepid.txt): tracking only the STEP in
// which the running extremum improved (FMUL + FMNMX per element and one compare + 2 selects per
// accumulator and step: 2.25 instead of 4 instructions per element) and re-reading that step's U
// entries at the end of the item to find the winning entry.  Bit-exact, but the re-read costs
// VEC x U scattered 4-byte loads per lane and item = +25 % L2 sectors on a kernel that is bound by
// L2 gather bandwidth, not by issue slots: Reddit-shape K=128 max 5.4 -> 6.3 ms (lean256),
// 5.1 -> 5.3-5.6 ms (lean128).
#ifndef ISPLIB_LEANMAX_MINB
// CTAs/SM the 32-byte max/min body is compiled for.  8 (64 registers, ~50 bytes of spills outside
// the gather loop, 32 warps/SM) beats 6 (80 registers, 24 warps/SM) and 7 on Reddit-shape:
// K=128 4.56 vs 4.93 / 4.75 ms, K=256 9.28 vs 9.67 / 9.41 ms (profiles/r2_kbench_max_minb.txt)
#define ISPLIB_LEANMAX_MINB 8
#endif
template <int OP, int VEC, int G, bool RAGGED, bool NOVAL>
__global__ void __launch_bounds__(128, VEC == 8 ? (OP == OP_SUM ? 8 : ISPLIB_LEANMAX_MINB) : (OP == OP_SUM ? 10 : 9))
spmm_lean_kernel(const __grid_constant__ SpmmParams p) {
    constexpr int U = 4;
    constexpr int NG = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int bx = blockIdx.x;
    if (p.gather.copy_ctas > 0) {
        // fused all-gather: the first CTAs of the grid push this rank's slice of x to the peers over NVLink
        if (bx < p.gather.copy_ctas) {
            if (blockIdx.y == 0) gather_push_role(p.gather);
            return;
        }
        bx -= p.gather.copy_ctas;
        if (p.gather.phase == 1) return;          // push-only launch (single-GPU emulation of the ranks)
        // tile mode: has this CTA's K tile of every peer's slice landed?  (before any warp may leave the CTA)
        if (p.gather.tile_vec4 > 0) gather_wait_for_tile(p.gather, blockIdx.y + p.tile_base);
    }
    const int item = bx * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= p.num_items) return;
    const int4 desc = __ldg(p.item_desc + item);
    const int eb = desc.y, ee = desc.z;
    const int g = lane / G;
    const bool lane_ok = !RAGGED || (lane % G) * VEC < p.tile_w;
    const int k0 = (blockIdx.y + p.tile_base) * p.tile_w + (lane_ok ? (lane % G) * VEC : 0);
    const char* const xlane = reinterpret_cast<const char*>(p.x + k0);
    const unsigned ldxb = (unsigned)p.ldx * 4u;
    const bool has_val = !NOVAL && (p.val != nullptr);
    // -0.0f the compiler cannot see through (kp is never negative): see track_better
    const float neg_zero = __int_as_float((int)(0x80000000u ^ (unsigned)(p.kp < 0)));

    float acc[1][VEC];
    int arg[1][VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { acc[0][v] = init_value<OP>(); arg[0][v] = kNoArg; }

#if ISPLIB_LEAN_PREFETCH
    unsigned c_next = 0;
    float a_next = 1.f;
    if (eb + lane < ee) {
        c_next = (unsigned)__ldcs(p.col + eb + lane);
        if (has_val) a_next = __ldcs(p.val + eb + lane);
    }
#endif
    // fused all-gather, owner mode: have the rows of this item's arrival group landed?  (checked here,
    // with the first index chunk already in flight, so the flag's round trip hides behind it)
    if (p.gather.copy_ctas > 0 && p.gather.tile_vec4 <= 0) gather_wait_for_item(p.gather, item, lane);
    for (int e0 = eb; e0 < ee; e0 += 32) {
        const int cnt = min(32, ee - e0);
#if ISPLIB_LEAN_PREFETCH
        const unsigned c = c_next;
        const float a = a_next;
        if (e0 + 32 + lane < ee) {
            c_next = (unsigned)__ldcs(p.col + e0 + 32 + lane);
            if (has_val) a_next = __ldcs(p.val + e0 + 32 + lane);
        }
#else
        unsigned c = 0;
        float a = 0.f;
        if (lane < cnt) {
            c = (unsigned)__ldcs(p.col + e0 + lane);
            a = has_val ? __ldcs(p.val + e0 + lane) : 1.f;
        }
#endif
        if (cnt == 32) {
#pragma unroll 1
            for (int t = 0; t < 32; t += NG * U) {
                float xv[U][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned cc = __shfl_sync(FULL, c, t + u * NG + g);
                    load_vec<VEC>(reinterpret_cast<const float*>(xlane + (unsigned long long)cc * ldxb), xv[u]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float aa = 1.f;
                    if constexpr (!NOVAL) aa = __shfl_sync(FULL, a, t + u * NG + g);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if constexpr (OP == OP_SUM) {
                            acc[0][v] = fmaf(aa, xv[u][v], acc[0][v]);
                        } else if constexpr (NOVAL) {
                            const float tt = xv[u][v];
                            if (better<OP>(tt, acc[0][v])) { acc[0][v] = tt; arg[0][v] = e0 + t + u * NG + g; }
                        } else {
                            track_better<OP>(aa, xv[u][v], e0 + t + u * NG + g, neg_zero, acc[0][v], arg[0][v]);
                        }
                    }
                }
            }
        } else {
            // last, short chunk of the segment: one step at a time (a U-wide predicated tail was
            // measured 5-10 % slower overall: it perturbs the register allocation of the main loop)
            for (int t = 0; t < cnt; t += NG) {
                const int idx = t + g;
                const unsigned cc = __shfl_sync(FULL, c, idx & 31);
                float aa = 1.f;
                if constexpr (!NOVAL) aa = __shfl_sync(FULL, a, idx & 31);
                if (idx < cnt) {
                    float xv[VEC];
                    load_vec<VEC>(reinterpret_cast<const float*>(xlane + (unsigned long long)cc * ldxb), xv);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if constexpr (OP == OP_SUM) {
                            acc[0][v] = fmaf(aa, xv[v], acc[0][v]);
                        } else {
                            const float tt = NOVAL ? xv[v] : __fmul_rn(aa, xv[v]);
                            if (better<OP>(tt, acc[0][v])) { acc[0][v] = tt; arg[0][v] = e0 + idx; }
                        }
                    }
                }
            }
        }
    }
    const int koff[1] = {k0};
    const bool kok[1] = {lane_ok};
    finish_item<OP, VEC, G, 1>(p, lane, desc.x, eb, ee, desc.w, koff, kok, acc, arg);
}

template <int OP, int VEC, bool NOVAL>
static inline SegKernel pick_lean_g(int g, bool ragged) {
    switch (g) {
        case 4: return ragged ? spmm_lean_kernel<OP, VEC, 4, true, NOVAL> : spmm_lean_kernel<OP, VEC, 4, false, NOVAL>;
        case 8: return ragged ? spmm_lean_kernel<OP, VEC, 8, true, NOVAL> : spmm_lean_kernel<OP, VEC, 8, false, NOVAL>;
        case 16: return ragged ? spmm_lean_kernel<OP, VEC, 16, true, NOVAL> : spmm_lean_kernel<OP, VEC, 16, false, NOVAL>;
        case 32: return ragged ? spmm_lean_kernel<OP, VEC, 32, true, NOVAL> : spmm_lean_kernel<OP, VEC, 32, false, NOVAL>;
        default: return nullptr;
    }
}

template <int OP, int VEC>
static inline SegKernel pick_lean(int g, bool ragged, bool noval) {
    if constexpr (OP != OP_SUM) {
        if (noval) return pick_lean_g<OP, VEC, true>(g, ragged);
    }
    return pick_lean_g<OP, VEC, false>(g, ragged);
}

struct TileShape { int vec, g, lpl, tile_w, ntiles; };

template <int OP, int VEC, int G, int LPL>
static inline SegKernel pick_u(int u, bool partial) {
    constexpr int NG = 32 / G;
    // keep (entries per step) * U <= 32 and at most 64 staged floats per lane
    if constexpr (LPL * VEC * 8 <= 64)
        if (u >= 8 && NG * 8 <= 32)
            return partial ? spmm_seg_kernel<OP, VEC, G, LPL, 8, true> : spmm_seg_kernel<OP, VEC, G, LPL, 8, false>;
    if (u <= 2)
        return partial ? spmm_seg_kernel<OP, VEC, G, LPL, 2, true> : spmm_seg_kernel<OP, VEC, G, LPL, 2, false>;
    if (NG * 4 <= 32)
        return partial ? spmm_seg_kernel<OP, VEC, G, LPL, 4, true> : spmm_seg_kernel<OP, VEC, G, LPL, 4, false>;
    return nullptr;
}

template <int OP, int G, int LPL>
static inline SegKernel pick_u8(int u, bool partial) {   // 32-byte vectors: U = 2 or 4 gathers in flight
    if (u <= 2)
        return partial ? spmm_seg_kernel<OP, 8, G, LPL, 2, true> : spmm_seg_kernel<OP, 8, G, LPL, 2, false>;
    return partial ? spmm_seg_kernel<OP, 8, G, LPL, 4, true> : spmm_seg_kernel<OP, 8, G, LPL, 4, false>;
}

template <int OP>
static inline SegKernel pick_kernel(const TileShape& t, int u, bool partial) {
    if (t.vec == 8) {
        if (t.g == 4) return pick_u8<OP, 4, 1>(u, partial);
        if (t.g == 8) return pick_u8<OP, 8, 1>(u, partial);
        if (t.g == 16) return pick_u8<OP, 16, 1>(u, partial);
        if (t.g == 32 && t.lpl == 1) return pick_u8<OP, 32, 1>(u, partial);
        if (t.g == 32 && t.lpl == 2) return pick_u8<OP, 32, 2>(u <= 2 ? 2 : 2, partial);
        return nullptr;
    }
    if (t.vec == 4) {
        if (t.g == 8 && t.lpl == 1) return pick_u<OP, 4, 8, 1>(u, partial);
        if (t.g == 16 && t.lpl == 1) return pick_u<OP, 4, 16, 1>(u, partial);
        if (t.g == 32 && t.lpl == 1) return pick_u<OP, 4, 32, 1>(u, partial);
        if (t.g == 32 && t.lpl == 2) return pick_u<OP, 4, 32, 2>(u, partial);
        if (t.g == 32 && t.lpl == 4) return pick_u<OP, 4, 32, 4>(u, partial);
    } else {
        if (t.lpl == 1) return pick_u<OP, 1, 32, 1>(u, partial);
        if (t.lpl == 2) return pick_u<OP, 1, 32, 2>(u, partial);
        if (t.lpl == 4) return pick_u<OP, 1, 32, 4>(u, partial);
    }
    return nullptr;
}

template <int OP, int G, int LPL>
static inline SegKernel pick_bulk_stages(int stages) {
    (void)stages;
    return spmm_bulk_kernel<OP, G, LPL, 3>;
}

template <int OP>
static inline SegKernel pick_bulk_kernel(const TileShape& t, int stages) {
    if (t.vec != 4) return nullptr;
    if (t.g == 8 && t.lpl == 1) return pick_bulk_stages<OP, 8, 1>(stages);
    if (t.g == 16 && t.lpl == 1) return pick_bulk_stages<OP, 16, 1>(stages);
    if (t.g == 32 && t.lpl == 1) return pick_bulk_stages<OP, 32, 1>(stages);
    if (t.g == 32 && t.lpl == 2) return pick_bulk_stages<OP, 32, 2>(stages);
    if (t.g == 32 && t.lpl == 4) return pick_bulk_stages<OP, 32, 4>(stages);
    return nullptr;
}


}  // namespace isplib
