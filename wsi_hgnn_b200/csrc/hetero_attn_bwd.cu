// Heterogeneous edge attention backward (kernel K3): gradients of wsi_hetero_attn_fwd (HEAT scoring) with respect to
// K, V, Q and the e_linear scalars - what DGL's GSDDMM / EdgeSoftmax / GSpMM backward functions compute for the
// reference's loss.backward() (trainer/train_gnn.py:68-71 through models/HEATNet4.py:103-119).
//
// One warp per dst row, lanes span D in the lane-grouped column order (wsi_head_perm).  Nothing is saved by the forward:
// per (row, relation) segment the scores are recomputed
//   AB m = max_e s_e, Z = sum_e exp(s_e - m), delta = sum_e a_e <g, V[src]>  (one online sweep over K[src], V[src])
//   C  ds_e = a_e (<g, V[src]> - delta)                                    (reads K[src], V[src] - L1/L2 hits)
//      dQ[row] += ds_e c_e K[src];  dK[src] += ds_e c_e q;  dV[src] += a_e g;   d c_e = ds_e <q, K[src]>
// (segments of one or two edges, the common case, keep their rows in registers and read them once)
// with g = dAgg[row] / R_t and c_e = (w sim_e + b) / sqrt(d_k).  dQ rows are owned by their warp; dK / dV rows are
// shared between destinations.  Two ways to accumulate them:
//   * TWO PASSES, no atomics (given the transposed graph): this kernel writes only the per-(edge, head) coefficients
//     cK[e,h] = ds_e c_e and cV[e,h] = a_e / R_t; attn_bwd_src_kernel then walks the SOURCE-major edge list - every
//     source row u owns dK[u] = sum_e cK[e] q[dst e], dV[u] = sum_e cV[e] dAgg[dst e]: the forward's gather pattern
//     (k-NN out-degree is constant, so it is perfectly balanced), deterministic, rows written once;
//   * one pass with 16-byte vector atomics (fp32 red.add; measured 3.4x the forward: bound by L2 atomic throughput).
#include "common.cuh"

namespace {

constexpr int WARPS = 4;
constexpr unsigned FULL = 0xffffffffu;

// sum over the G = 32 / H lanes of a head (G a power of two): five warp-uniform predicated steps - as a run-time loop this
// was 8 instructions per step and edge (loop counter, divergence check, branch) around one SHFL + one FADD
__device__ __forceinline__ float head_reduce(float d, int G) {
  if (G >= 32) d += __shfl_xor_sync(0xffffffffu, d, 16);
  if (G >= 16) d += __shfl_xor_sync(0xffffffffu, d, 8);
  if (G >= 8) d += __shfl_xor_sync(0xffffffffu, d, 4);
  if (G >= 4) d += __shfl_xor_sync(0xffffffffu, d, 2);
  if (G >= 2) d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}


struct BwdArgs {
  const float* K; int64_t ldk;
  const float* V; int64_t ldv;
  const float* Q; int64_t ldq;
  const int* rowptr; const int* e_src; const float* e_sim; const uint8_t* e_rel;
  const float* inv_r; const float* e_w; const float* e_b;
  const float* dAgg; int64_t ldg;
  float* dK; int64_t lddk;
  float* dV; int64_t lddv;
  float* dQ; int64_t lddq;
  float* d_e;                 // [2]: d e_linear.weight, d e_linear.bias (accumulated)
  const int* order;           // optional [n_rows]: processing order of the rows (largest in-degree first)
  int n_rows, D, H;
  float inv_sqrt_dk;
  float* coef;                // two-pass mode: [E, 2, H] per-edge (cK | cV); nullptr = atomics
  // second pass (source-major edge list)
  const int* t_ptr; const int* t_eid; const int* t_dst; int n_src;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void red_add4(float* p, float4 v) {
#if __CUDA_ARCH__ >= 900
  atomicAdd(reinterpret_cast<float4*>(p), v);
#else
  atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
#endif
}

template <int NV>
__device__ __forceinline__ float head_dot(const float4* a, const float4* b, int G) {
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    d0 = fmaf(a[i].x, b[i].x, d0); d1 = fmaf(a[i].y, b[i].y, d1);
    d0 = fmaf(a[i].z, b[i].z, d0); d1 = fmaf(a[i].w, b[i].w, d1);
  }
  float d = d0 + d1;
  d = head_reduce(d, G);
  return d;
}

// per-edge gradient contributions of one edge whose K / V rows are in registers
template <int NV>
__device__ __forceinline__ void bwd_edge(const BwdArgs& a, int lane, int e, int src, float dsc, float at, float invr,
                                         const float4* kk, const float4* q, const float4* g, float4* dq) {
  if (a.coef) {                                            // two-pass mode: only the coefficients leave this kernel
    const int G = 32 / a.H;
    if (lane % G == 0) {
      float* c = a.coef + (int64_t)e * 2 * a.H + lane / G;
      c[0] = dsc;
      c[a.H] = at * invr;                                  // (g already carries 1 / R_t; the second pass reads raw dAgg)
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      dq[i].x = fmaf(dsc, kk[i].x, dq[i].x); dq[i].y = fmaf(dsc, kk[i].y, dq[i].y);
      dq[i].z = fmaf(dsc, kk[i].z, dq[i].z); dq[i].w = fmaf(dsc, kk[i].w, dq[i].w);
    }
    return;
  }
  float* dkr = a.dK + (int64_t)src * a.lddk;
  float* dvr = a.dV + (int64_t)src * a.lddv;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dq[i].x = fmaf(dsc, kk[i].x, dq[i].x); dq[i].y = fmaf(dsc, kk[i].y, dq[i].y);
    dq[i].z = fmaf(dsc, kk[i].z, dq[i].z); dq[i].w = fmaf(dsc, kk[i].w, dq[i].w);
    red_add4(dkr + (i * 32 + lane) * 4, make_float4(dsc * q[i].x, dsc * q[i].y, dsc * q[i].z, dsc * q[i].w));
    red_add4(dvr + (i * 32 + lane) * 4, make_float4(at * g[i].x, at * g[i].y, at * g[i].z, at * g[i].w));
  }
}

// Second pass of the two-pass mode: one warp per SOURCE row, two out-edges in flight.
template <int NV>
__global__ void __launch_bounds__(WARPS * 32) attn_bwd_src_kernel(BwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int head = lane / (32 / a.H);
  const int u = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (u >= a.n_src) return;
  const int beg = __ldg(a.t_ptr + u), end = __ldg(a.t_ptr + u + 1);
  float4 ak[NV], av[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ak[i] = make_float4(0.f, 0.f, 0.f, 0.f); av[i] = ak[i]; }
  for (int j = beg; j < end; j += 2) {
    const bool two = j + 1 < end;
    const int e0 = __ldg(a.t_eid + j), e1 = two ? __ldg(a.t_eid + j + 1) : e0;
    const int v0 = __ldg(a.t_dst + j), v1 = two ? __ldg(a.t_dst + j + 1) : v0;
    const float ck0 = __ldg(a.coef + (int64_t)e0 * 2 * a.H + head), cv0 = __ldg(a.coef + (int64_t)e0 * 2 * a.H + a.H + head);
    const float ck1 = two ? __ldg(a.coef + (int64_t)e1 * 2 * a.H + head) : 0.f;
    const float cv1 = two ? __ldg(a.coef + (int64_t)e1 * 2 * a.H + a.H + head) : 0.f;
    float4 q0[NV], q1[NV], g0[NV], g1[NV];
    const float* qr0 = a.Q + (int64_t)v0 * a.ldq; const float* qr1 = a.Q + (int64_t)v1 * a.ldq;
    const float* gr0 = a.dAgg + (int64_t)v0 * a.ldg; const float* gr1 = a.dAgg + (int64_t)v1 * a.ldg;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      q0[i] = ld4(qr0 + (i * 32 + lane) * 4); q1[i] = ld4(qr1 + (i * 32 + lane) * 4);
      g0[i] = ld4(gr0 + (i * 32 + lane) * 4); g1[i] = ld4(gr1 + (i * 32 + lane) * 4);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      ak[i].x = fmaf(ck0, q0[i].x, fmaf(ck1, q1[i].x, ak[i].x)); ak[i].y = fmaf(ck0, q0[i].y, fmaf(ck1, q1[i].y, ak[i].y));
      ak[i].z = fmaf(ck0, q0[i].z, fmaf(ck1, q1[i].z, ak[i].z)); ak[i].w = fmaf(ck0, q0[i].w, fmaf(ck1, q1[i].w, ak[i].w));
      av[i].x = fmaf(cv0, g0[i].x, fmaf(cv1, g1[i].x, av[i].x)); av[i].y = fmaf(cv0, g0[i].y, fmaf(cv1, g1[i].y, av[i].y));
      av[i].z = fmaf(cv0, g0[i].z, fmaf(cv1, g1[i].z, av[i].z)); av[i].w = fmaf(cv0, g0[i].w, fmaf(cv1, g1[i].w, av[i].w));
    }
  }
  float* dkr = a.dK + (int64_t)u * a.lddk;
  float* dvr = a.dV + (int64_t)u * a.lddv;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    *reinterpret_cast<float4*>(dkr + (i * 32 + lane) * 4) = ak[i];
    *reinterpret_cast<float4*>(dvr + (i * 32 + lane) * 4) = av[i];
  }
}

// One dst row per warp (rows are dealt out largest first through `order`, one row per warp, so the block scheduler
// balances the heavy-tailed in-degrees).  Per (row, relation) segment:
//   n <= 2 edges (the common case of a k-NN graph spread over up to 2 T^2 relations): K and V rows are read ONCE into
//       registers, everything else is arithmetic on them;
//   longer segments: pass AB (online softmax statistics and delta together; K and V rows, two edges in flight),
//       pass C (gradients; K and V rows again, two edges in flight).
template <int NV>
__global__ void __launch_bounds__(WARPS * 32, NV <= 4 ? 3 : 1) attn_bwd_kernel(BwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int G = 32 / a.H;
  const int n_warps = gridDim.x * WARPS;
  const float ew = __ldg(a.e_w), eb = __ldg(a.e_b);
  float dw_acc = 0.f, db_acc = 0.f;

  for (int idx = blockIdx.x * WARPS + (threadIdx.x >> 5); idx < a.n_rows; idx += n_warps) {
    const int row = a.order ? __ldg(a.order + idx) : idx;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const float invr = __ldg(a.inv_r + row);
    float4 dq[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) dq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (invr != 0.f && end > beg) {
      float4 q[NV], g[NV];
      const float* qr = a.Q + (int64_t)row * a.ldq;
      const float* gr = a.dAgg + (int64_t)row * a.ldg;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        q[i] = ld4(qr + (i * 32 + lane) * 4);
        g[i] = ld4(gr + (i * 32 + lane) * 4);
        g[i].x *= invr; g[i].y *= invr; g[i].z *= invr; g[i].w *= invr;
      }
      int seg_beg = beg;
      while (seg_beg < end) {
        // segment = run of edges of one relation; its ids / sims / relation bytes are read lane-parallel, 32 at a time
        const int rel = __ldg(a.e_rel + seg_beg);
        int seg_end = seg_beg;
        for (int w0 = seg_beg; w0 < end; w0 += 32) {
          const int r = (w0 + lane < end) ? (int)__ldg(a.e_rel + w0 + lane) : -1;
          const unsigned diff = __ballot_sync(FULL, r != rel);
          if (diff) { seg_end = w0 + __ffs(diff) - 1; break; }
          seg_end = min(end, w0 + 32);
        }
        const int n = seg_end - seg_beg;
        if (n <= 2) {
          // ---- register path: both rows of K and V read once
          const int s0 = __ldg(a.e_src + seg_beg), s1 = n > 1 ? __ldg(a.e_src + seg_beg + 1) : s0;
          const float sim0 = __ldg(a.e_sim + seg_beg), sim1 = n > 1 ? __ldg(a.e_sim + seg_beg + 1) : 0.f;
          float4 k0[NV], k1[NV], v0[NV], v1[NV];
          const float* kr0 = a.K + (int64_t)s0 * a.ldk; const float* kr1 = a.K + (int64_t)s1 * a.ldk;
          const float* vr0 = a.V + (int64_t)s0 * a.ldv; const float* vr1 = a.V + (int64_t)s1 * a.ldv;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            k0[i] = ld4(kr0 + (i * 32 + lane) * 4); k1[i] = ld4(kr1 + (i * 32 + lane) * 4);
            v0[i] = ld4(vr0 + (i * 32 + lane) * 4); v1[i] = ld4(vr1 + (i * 32 + lane) * 4);
          }
          const float c0 = fmaf(ew, sim0, eb) * a.inv_sqrt_dk, c1 = fmaf(ew, sim1, eb) * a.inv_sqrt_dk;
          const float d0 = head_dot<NV>(q, k0, G), d1 = head_dot<NV>(q, k1, G);
          const float sc0 = d0 * c0, sc1 = n > 1 ? d1 * c1 : -INFINITY;
          const float m = fmaxf(sc0, sc1);
          const float p0 = __expf(sc0 - m), p1 = __expf(sc1 - m);        // exp(-inf) = 0 for the absent edge
          const float inv_z = 1.f / (p0 + p1);
          const float a0 = p0 * inv_z, a1 = p1 * inv_z;
          const float t0 = head_dot<NV>(g, v0, G), t1 = head_dot<NV>(g, v1, G);
          const float delta = a0 * t0 + a1 * t1;
          const float ds0 = a0 * (t0 - delta), ds1 = a1 * (t1 - delta);
          bwd_edge<NV>(a, lane, seg_beg, s0, ds0 * c0, a0, invr, k0, q, g, dq);
          if (n > 1) bwd_edge<NV>(a, lane, seg_beg + 1, s1, ds1 * c1, a1, invr, k1, q, g, dq);
          if (lane % G == 0) {                            // one lane per head carries the head's d c_e
            const float dc0 = ds0 * d0 * a.inv_sqrt_dk, dc1 = n > 1 ? ds1 * d1 * a.inv_sqrt_dk : 0.f;
            dw_acc = fmaf(dc0, sim0, fmaf(dc1, sim1, dw_acc));
            db_acc += dc0 + dc1;
          }
        } else if (a.coef) {
          // ---- ONE sweep over K[src], V[src] (two-pass mode; measured round 2: the second sweep of the AB / C scheme
          // below missed L1 / L2 for most rows at training batch sizes - 8.2 GB of DRAM reads against 6.0 GB
          // algorithmic).  Everything that needs the final softmax statistics is kept in a form that can be rescaled:
          //   dQ_seg = sum_e ds_e c_e K_e,  ds_e = a_e (t_e - delta),  a_e = p_e / Z,  t_e = <g, V_e>
          //          = (A1 - delta A2) / Z   with  A1 = sum_e p_e t_e c_e K_e,  A2 = sum_e p_e c_e K_e
          // (online-softmax accumulators, rescaled whenever the running maximum moves), and the per-(edge, head) raw
          // values d_e = <q, K_e>, t_e are parked in the coefficient array and turned into cK / cV by a lane-parallel
          // fix-up once m, Z, delta are known.
          float m = -INFINITY, z = 0.f, num = 0.f;
          float4 a1[NV], a2[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i) { a1[i] = make_float4(0.f, 0.f, 0.f, 0.f); a2[i] = a1[i]; }
          for (int e = seg_beg; e < seg_end; e += 2) {
            const bool two = e + 1 < seg_end;
            const int s0 = __ldg(a.e_src + e), s1 = two ? __ldg(a.e_src + e + 1) : s0;
            const float c0 = fmaf(ew, __ldg(a.e_sim + e), eb) * a.inv_sqrt_dk;
            const float c1 = two ? fmaf(ew, __ldg(a.e_sim + e + 1), eb) * a.inv_sqrt_dk : 0.f;
            float4 k0[NV], k1[NV], v0[NV], v1[NV];
            const float* kr0 = a.K + (int64_t)s0 * a.ldk; const float* kr1 = a.K + (int64_t)s1 * a.ldk;
            const float* vr0 = a.V + (int64_t)s0 * a.ldv; const float* vr1 = a.V + (int64_t)s1 * a.ldv;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
              k0[i] = ld4(kr0 + (i * 32 + lane) * 4); k1[i] = ld4(kr1 + (i * 32 + lane) * 4);
              v0[i] = ld4(vr0 + (i * 32 + lane) * 4); v1[i] = ld4(vr1 + (i * 32 + lane) * 4);
            }
            const float d0 = head_dot<NV>(q, k0, G), d1 = head_dot<NV>(q, k1, G);
            const float sc0 = d0 * c0, sc1 = two ? d1 * c1 : -INFINITY;
            const float t0 = head_dot<NV>(g, v0, G), t1 = head_dot<NV>(g, v1, G);
            const float mn = fmaxf(m, fmaxf(sc0, sc1));
            const float corr = __expf(m - mn), p0 = __expf(sc0 - mn), p1 = __expf(sc1 - mn);   // exp(-inf) = 0
            z = fmaf(z, corr, p0 + p1);
            num = fmaf(num, corr, fmaf(p0, t0, p1 * t1));
            m = mn;
            const float w20 = p0 * c0, w21 = p1 * c1, w10 = w20 * t0, w11 = w21 * t1;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
              a1[i].x = fmaf(a1[i].x, corr, fmaf(w10, k0[i].x, w11 * k1[i].x)); a1[i].y = fmaf(a1[i].y, corr, fmaf(w10, k0[i].y, w11 * k1[i].y));
              a1[i].z = fmaf(a1[i].z, corr, fmaf(w10, k0[i].z, w11 * k1[i].z)); a1[i].w = fmaf(a1[i].w, corr, fmaf(w10, k0[i].w, w11 * k1[i].w));
              a2[i].x = fmaf(a2[i].x, corr, fmaf(w20, k0[i].x, w21 * k1[i].x)); a2[i].y = fmaf(a2[i].y, corr, fmaf(w20, k0[i].y, w21 * k1[i].y));
              a2[i].z = fmaf(a2[i].z, corr, fmaf(w20, k0[i].z, w21 * k1[i].z)); a2[i].w = fmaf(a2[i].w, corr, fmaf(w20, k0[i].w, w21 * k1[i].w));
            }
            if (lane % G == 0) {                            // park the raw per-(edge, head) values
              float* c = a.coef + (int64_t)e * 2 * a.H + lane / G;
              c[0] = d0; c[a.H] = t0;
              if (two) { c[2 * a.H] = d1; c[3 * a.H] = t1; }
            }
          }
          const float inv_z = 1.f / z;
          const float delta = num * inv_z;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            dq[i].x = fmaf(fmaf(-delta, a2[i].x, a1[i].x), inv_z, dq[i].x); dq[i].y = fmaf(fmaf(-delta, a2[i].y, a1[i].y), inv_z, dq[i].y);
            dq[i].z = fmaf(fmaf(-delta, a2[i].z, a1[i].z), inv_z, dq[i].z); dq[i].w = fmaf(fmaf(-delta, a2[i].w, a1[i].w), inv_z, dq[i].w);
          }
          // ---- fix-up: every lane takes (edge, head) pairs; the head's statistics come from a lane of its group
          __syncwarp();
          const int pairs = n * a.H;
          for (int p0i = 0; p0i < pairs; p0i += 32) {
            const int p = p0i + lane;
            const bool on = p < pairs;
            const int h = on ? p % a.H : 0, e = seg_beg + (on ? p / a.H : 0);
            const float mh = __shfl_sync(FULL, m, h * G), izh = __shfl_sync(FULL, inv_z, h * G), dh = __shfl_sync(FULL, delta, h * G);
            if (on) {
              volatile float* c = a.coef + (int64_t)e * 2 * a.H + h;
              const float d = c[0], t = c[a.H];
              const float sim = __ldg(a.e_sim + e);
              const float ce = fmaf(ew, sim, eb) * a.inv_sqrt_dk;
              const float at = __expf(d * ce - mh) * izh;
              const float ds = at * (t - dh);
              c[0] = ds * ce;
              c[a.H] = at * invr;
              const float dc = ds * d * a.inv_sqrt_dk;
              dw_acc = fmaf(dc, sim, dw_acc);
              db_acc += dc;
            }
          }
        } else {
          // ---- pass AB: m, Z and delta in one sweep (online softmax), two edges in flight
          float m = -INFINITY, z = 0.f, num = 0.f;
          for (int e = seg_beg; e < seg_end; e += 2) {
            const bool two = e + 1 < seg_end;
            const int s0 = __ldg(a.e_src + e), s1 = two ? __ldg(a.e_src + e + 1) : s0;
            const float c0 = fmaf(ew, __ldg(a.e_sim + e), eb) * a.inv_sqrt_dk;
            const float c1 = two ? fmaf(ew, __ldg(a.e_sim + e + 1), eb) * a.inv_sqrt_dk : 0.f;
            float4 k0[NV], k1[NV], v0[NV], v1[NV];
            const float* kr0 = a.K + (int64_t)s0 * a.ldk; const float* kr1 = a.K + (int64_t)s1 * a.ldk;
            const float* vr0 = a.V + (int64_t)s0 * a.ldv; const float* vr1 = a.V + (int64_t)s1 * a.ldv;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
              k0[i] = ld4(kr0 + (i * 32 + lane) * 4); k1[i] = ld4(kr1 + (i * 32 + lane) * 4);
              v0[i] = ld4(vr0 + (i * 32 + lane) * 4); v1[i] = ld4(vr1 + (i * 32 + lane) * 4);
            }
            const float sc0 = head_dot<NV>(q, k0, G) * c0, sc1 = two ? head_dot<NV>(q, k1, G) * c1 : -INFINITY;
            const float t0 = head_dot<NV>(g, v0, G), t1 = head_dot<NV>(g, v1, G);
            const float mn = fmaxf(m, fmaxf(sc0, sc1));
            const float corr = __expf(m - mn), p0 = __expf(sc0 - mn), p1 = __expf(sc1 - mn);
            z = fmaf(z, corr, p0 + p1);
            num = fmaf(num, corr, fmaf(p0, t0, p1 * t1));
            m = mn;
          }
          const float inv_z = 1.f / z;
          const float delta = num * inv_z;
          // ---- pass C: the gradients, two edges in flight
          for (int e = seg_beg; e < seg_end; e += 2) {
            const bool two = e + 1 < seg_end;
            const int s0 = __ldg(a.e_src + e), s1 = two ? __ldg(a.e_src + e + 1) : s0;
            const float sim0 = __ldg(a.e_sim + e), sim1 = two ? __ldg(a.e_sim + e + 1) : 0.f;
            const float c0 = fmaf(ew, sim0, eb) * a.inv_sqrt_dk, c1 = fmaf(ew, sim1, eb) * a.inv_sqrt_dk;
            float4 k0[NV], k1[NV], v0[NV], v1[NV];
            const float* kr0 = a.K + (int64_t)s0 * a.ldk; const float* kr1 = a.K + (int64_t)s1 * a.ldk;
            const float* vr0 = a.V + (int64_t)s0 * a.ldv; const float* vr1 = a.V + (int64_t)s1 * a.ldv;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
              k0[i] = ld4(kr0 + (i * 32 + lane) * 4); k1[i] = ld4(kr1 + (i * 32 + lane) * 4);
              v0[i] = ld4(vr0 + (i * 32 + lane) * 4); v1[i] = ld4(vr1 + (i * 32 + lane) * 4);
            }
            const float d0 = head_dot<NV>(q, k0, G), d1 = head_dot<NV>(q, k1, G);
            const float a0 = __expf(d0 * c0 - m) * inv_z, a1 = two ? __expf(d1 * c1 - m) * inv_z : 0.f;
            const float ds0 = a0 * (head_dot<NV>(g, v0, G) - delta), ds1 = a1 * (head_dot<NV>(g, v1, G) - delta);
            bwd_edge<NV>(a, lane, e, s0, ds0 * c0, a0, invr, k0, q, g, dq);
            if (two) bwd_edge<NV>(a, lane, e + 1, s1, ds1 * c1, a1, invr, k1, q, g, dq);
            if (lane % G == 0) {
              const float dc0 = ds0 * d0 * a.inv_sqrt_dk, dc1 = two ? ds1 * d1 * a.inv_sqrt_dk : 0.f;
              dw_acc = fmaf(dc0, sim0, fmaf(dc1, sim1, dw_acc));
              db_acc += dc0 + dc1;
            }
          }
        }
        seg_beg = seg_end;
      }
    }
    float* o = a.dQ + (int64_t)row * a.lddq;
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(o + (i * 32 + lane) * 4) = dq[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dw_acc += __shfl_xor_sync(FULL, dw_acc, o);
    db_acc += __shfl_xor_sync(FULL, db_acc, o);
  }
  if (lane == 0 && (dw_acc != 0.f || db_acc != 0.f)) { atomicAdd(a.d_e, dw_acc); atomicAdd(a.d_e + 1, db_acc); }
}

}  // namespace

// dK, dV: ACCUMULATED into (zero them first) in the one-pass mode, WRITTEN (all n_src rows) in the two-pass mode;
// dQ: written; d_e [2]: accumulated.
extern "C" int wsi_hetero_attn_bwd(const float* k, int64_t ldk, const float* v, int64_t ldv, const float* q, int64_t ldq,
                                   const int32_t* rowptr, const int32_t* e_src, const float* e_sim, const uint8_t* e_rel,
                                   const float* node_inv_r, const float* e_w, const float* e_b, int64_t n_rows, int D,
                                   int H, const float* d_agg, int64_t ldg, float* dk, int64_t lddk, float* dv,
                                   int64_t lddv, float* dq, int64_t lddq, float* d_e, const int32_t* row_order,
                                   const int32_t* t_ptr, const int32_t* t_eid, const int32_t* t_dst, int64_t n_src,
                                   float* coef_ws, void* stream) {
  WSI_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "hetero_attn_bwd: bad n_rows");
  if (n_rows == 0) return WSI_OK;
  WSI_CHECK_ARG(k && v && q && rowptr && node_inv_r && e_w && e_b && d_agg && dk && dv && dq && d_e,
                "hetero_attn_bwd: null pointer");
  if (!(D % 128 == 0 && D <= 1024 && H >= 1 && H <= 32 && (H & (H - 1)) == 0 && D % H == 0)) {
    wsi_set_error("hetero_attn_bwd: needs the lane-grouped layout (D %% 128 == 0, D <= 1024, H a power of two <= 32), got D=%d H=%d", D, H);
    return WSI_ERR_UNSUPPORTED;
  }
  WSI_CHECK_ARG(ldk % 4 == 0 && ldv % 4 == 0 && ldq % 4 == 0 && ldg % 4 == 0 && lddk % 4 == 0 && lddv % 4 == 0 && lddq % 4 == 0,
                "hetero_attn_bwd: row strides must be multiples of 4 floats");
  BwdArgs a{};
  a.K = k; a.ldk = ldk; a.V = v; a.ldv = ldv; a.Q = q; a.ldq = ldq;
  a.rowptr = rowptr; a.e_src = e_src; a.e_sim = e_sim; a.e_rel = e_rel; a.inv_r = node_inv_r; a.e_w = e_w; a.e_b = e_b;
  a.dAgg = d_agg; a.ldg = ldg; a.dK = dk; a.lddk = lddk; a.dV = dv; a.lddv = lddv; a.dQ = dq; a.lddq = lddq; a.d_e = d_e;
  a.n_rows = (int)n_rows; a.D = D; a.H = H; a.inv_sqrt_dk = 1.0f / sqrtf((float)(D / H));
  a.order = row_order;
  const bool two_pass = t_ptr != nullptr;
  WSI_CHECK_ARG(!two_pass || (t_eid && t_dst && coef_ws && n_src >= 0 && n_src < (1ll << 31)),
                "hetero_attn_bwd: the two-pass mode needs t_ptr, t_eid, t_dst, coef_ws and n_src");
  a.coef = two_pass ? coef_ws : nullptr;
  a.t_ptr = t_ptr; a.t_eid = t_eid; a.t_dst = t_dst; a.n_src = (int)n_src;
  const int blocks = (int)((n_rows + WARPS - 1) / WARPS);     // one row per warp: the block scheduler is the queue
  cudaStream_t st = wsi_stream(stream);
  switch (D / 128) {
#define CASE(NV) case NV: attn_bwd_kernel<NV><<<blocks, WARPS * 32, 0, st>>>(a); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  WSI_CHECK_LAUNCH();
  if (two_pass && n_src > 0) {
    const int sb = (int)((n_src + WARPS - 1) / WARPS);
    switch (D / 128) {
#define CASE(NV) case NV: attn_bwd_src_kernel<NV><<<sb, WARPS * 32, 0, st>>>(a); break;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    WSI_CHECK_LAUNCH();
  }
  return WSI_OK;
}
