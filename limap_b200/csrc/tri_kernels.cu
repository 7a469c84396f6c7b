// tri_kernels.cu — sm_100a kernels of the line-triangulation hot path.
//
//   tri_node_kernel : one CTA per 2D line ("node"). Fuses
//       BaseLineTriangulator::triangulateOneNode   (triangulation/base_line_triangulator.cc:161-337)
//       GlobalLineTriangulator::scoreOneNode       (triangulation/global_line_triangulator.cc:71-161)
//     so that the candidate set of a node lives only in shared memory: match rows are read once from
//     HBM, candidates are never written back unless debug_mode asks for them.
//   expand_rows / node_offsets : turn the per-(image, neighbour) match tables into node-major rows.
//   collect_edges / edge_weights : run_clustering's edge list and 3d scores
//       (global_line_triangulator.cc:234-291).
//
// Arithmetic is fp64 in this first (exact) path: every gate is a hard threshold on a transcendental
// and the contract is bit-exact candidate indices (see DESIGN.md "precision").
#include "tri_kernels.cuh"
#include <cstdio>

namespace lm {

static constexpr int kThreads = 128;
static constexpr int kWarps = kThreads / 32;
static constexpr int kCandDoubles = 17; // sx..ez(6) d(3) zs ze unc q(4) score
static constexpr int kCandBytes = kCandDoubles * 8 + 8;

size_t tri_smem_bytes(int cap) { return (size_t)cap * kCandBytes; }

struct Slab {
  double *sx, *sy, *sz, *ex, *ey, *ez, *dx, *dy, *dz, *zs, *ze, *unc, *q0, *q1, *q2, *q3, *score;
  uint32_t *ng, *row;
  LM_D void carve(char *base, int cap) {
    double *d = reinterpret_cast<double *>(base);
    sx = d; sy = sx + cap; sz = sy + cap; ex = sz + cap; ey = ex + cap; ez = ey + cap;
    dx = ez + cap; dy = dx + cap; dz = dy + cap; zs = dz + cap; ze = zs + cap; unc = ze + cap;
    q0 = unc + cap; q1 = q0 + cap; q2 = q1 + cap; q3 = q2 + cap; score = q3 + cap;
    ng = reinterpret_cast<uint32_t *>(score + cap);
    row = ng + cap;
  }
};

struct Cand {
  vec3<double> s, e;
  double zs, ze, unc;
};

LM_D double4 ld_seg(const double4 *p) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
  const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
LM_D double deg_from_cos_abs(double c) { return acos(fabs(c)) * consts<double>::rad2deg(); }

// Line3d::sensitivity (base/linebase.cc:100-107)
LM_D double sensitivity(const ViewD &v, vec3<double> Xs, vec3<double> Xe, vec3<double> dir) {
  vec2<double> ps = dehom(proj_h(v.P, Xs));
  vec2<double> pe = dehom(proj_h(v.P, Xe));
  vec2<double> mid = (ps + pe) * 0.5;
  vec3<double> d3 = normalized(mat3_mul_h(v.M, mid.x, mid.y));
  return 90.0 - deg_from_cos_abs(dot(dir, d3));
}

// triangulate_point (triangulation/functions.cc:100-117): mid-point method, 2x2 LDLT.
LM_D bool triangulate_point(const ViewD &v1, const ViewD &v2, vec3<double> n1e, vec3<double> n2e, vec3<double> C1,
                            vec3<double> C2, vec3<double> &out) {
  double a00 = dot(n1e, n1e), a01 = -dot(n1e, n2e), a10 = -dot(n2e, n1e), a11 = dot(n2e, n2e);
  double b0 = dot(n1e, C2 - C1), b1 = dot(n2e, C1 - C2);
  double r0, r1;
  if (a00 >= a11) {
    double l10 = a10 / a00, d1 = a11 - l10 * a01, y1 = b1 - l10 * b0;
    r1 = y1 / d1;
    r0 = (b0 - a01 * r1) / a00;
  } else {
    double l01 = a01 / a11, d0 = a00 - l01 * a10, y0 = b0 - l01 * b1;
    r0 = y0 / d0;
    r1 = (b1 - a10 * r0) / a11;
  }
  out = (n1e * r0 + C1 + n2e * r1 + C2) * 0.5;
  double z1 = v1.P[8] * out.x + v1.P[9] * out.y + v1.P[10] * out.z + v1.P[11];
  double z2 = v2.P[8] * out.x + v2.P[9] * out.y + v2.P[10] * out.z + v2.P[11];
  return !(z1 < consts<double>::eps() || z2 < consts<double>::eps());
}

// Per-node constants of the source line (computed redundantly by every thread: ~60 flops).
struct Src {
  double4 l1;
  vec3<double> w1s, w1e;     // M1 [p;1] (unnormalised world rays)
  vec3<double> ray1s, ray1e; // normalised
  vec3<double> C1;
  bool ok;
};

// One match row -> candidate. Steps follow triangulateOneNode "Step 3" (base_line_triangulator.cc:290-326).
LM_D bool gen_candidate(const TriParams &p, const ViewD &v1, const Src &src, uint32_t ngv, uint32_t ngl, Cand &c,
                        double4 &l2out) {
  const double4 l2 = ld_seg(&p.segs[p.line_off[ngv] + ngl]);
  l2out = l2;
  {
    double dx = l2.x - l2.z, dy = l2.y - l2.w;
    if (sqrt(dx * dx + dy * dy) <= p.min_length_2d) return false; // :177
  }
  if (p.disable_algebraic) return false;
  const ViewD &v2 = p.views[ngv];
  vec3<double> c2s = mat3_mul_h(v2.M, l2.x, l2.y);
  vec3<double> c2e = mat3_mul_h(v2.M, l2.z, l2.w);
  // getNormalDirection (functions.cc:28-35) + ray-plane angle tests (:292-302)
  vec3<double> n2 = normalized(cross(c2s, c2e));
  double angle_start = 90.0 - deg_from_cos_abs(dot(n2, src.ray1s));
  if (angle_start < p.line_tri_angle_threshold) return false;
  double angle_end = 90.0 - deg_from_cos_abs(dot(n2, src.ray1e));
  if (angle_end < p.line_tri_angle_threshold) return false;
  vec3<double> C2 = mk3(v2.C[0], v2.C[1], v2.C[2]);
  // compute_epipolar_IoU (functions.cc:76-98). F x1 = M2^T ((C1 - C2) x (M1 x1)) exactly
  // (F = K2^-T [t]x R2 R1^T K1^-1 with t = R2 (C1 - C2)); see DESIGN.md.
  {
    vec3<double> base = src.C1 - C2;
    vec3<double> coor_l2 = normalized(cross(mk3(l2.x, l2.y, 1.0), mk3(l2.z, l2.w, 1.0)));
    vec3<double> eps_ = normalized(mat3T_mul(v2.M, cross(base, src.w1s)));
    vec2<double> cs = dehom(cross(coor_l2, eps_));
    vec3<double> epe_ = normalized(mat3T_mul(v2.M, cross(base, src.w1e)));
    vec2<double> ce = dehom(cross(coor_l2, epe_));
    vec2<double> s2 = mk2(l2.x, l2.y), e2 = mk2(l2.z, l2.w);
    vec2<double> dir2 = normalized(e2 - s2);
    double len2 = norm(s2 - e2);
    double c1 = dot(cs - s2, dir2) / len2;
    double c2 = dot(ce - s2, dir2) / len2;
    if (c1 > c2) { double t = c1; c1 = c2; c2 = t; }
    double IoU = (smin(c2, 1.0) - smax(c1, 0.0)) / (smax(c2, 1.0) - smin(c1, 0.0));
    if (IoU < p.IoU_threshold) return false;
  }
  vec3<double> Xs, Xe;
  vec3<double> r2s = normalized(c2s), r2e = normalized(c2e);
  const double EPS = consts<double>::eps();
  if (!p.use_endpoints_triangulation) {
    // line_triangulation (functions.cc:194-233): plane-pair intersection
    vec3<double> B = C2 - src.C1;
    vec3<double> nb = mk3(-r2s.x, -r2s.y, -r2s.z), nc = mk3(-r2e.x, -r2e.y, -r2e.z);
    vec3<double> ls = solve3_cols(src.ray1s, nb, nc, B);
    Xs = src.ray1s * ls.x + src.C1;
    vec3<double> le = solve3_cols(src.ray1e, nb, nc, B);
    Xe = src.ray1e * le.x + src.C1;
    c.zs = v1.P[8] * Xs.x + v1.P[9] * Xs.y + v1.P[10] * Xs.z + v1.P[11];
    c.ze = v1.P[8] * Xe.x + v1.P[9] * Xe.y + v1.P[10] * Xe.z + v1.P[11];
    if (c.zs < EPS || c.ze < EPS) return false;
    double d21 = v2.P[8] * Xs.x + v2.P[9] * Xs.y + v2.P[10] * Xs.z + v2.P[11];
    double d22 = v2.P[8] * Xe.x + v2.P[9] * Xe.y + v2.P[10] * Xe.z + v2.P[11];
    if (d21 < EPS || d22 < EPS) return false;
    if (isnan(Xs.x) || isnan(Xe.x)) return false;
  } else {
    // triangulate_line_by_endpoints (functions.cc:172-190)
    if (!triangulate_point(v1, v2, src.ray1s, r2s, src.C1, C2, Xs)) return false;
    if (!triangulate_point(v1, v2, src.ray1e, r2e, src.C1, C2, Xe)) return false;
    c.zs = v1.P[8] * Xs.x + v1.P[9] * Xs.y + v1.P[10] * Xs.z + v1.P[11];
    c.ze = v1.P[8] * Xe.x + v1.P[9] * Xe.y + v1.P[10] * Xe.z + v1.P[11];
  }
  // sensitivity in both views (:315-317)
  vec3<double> dir = normalized(Xe - Xs);
  if (sensitivity(v1, Xs, Xe, dir) > p.sensitivity_threshold &&
      sensitivity(v2, Xs, Xe, dir) > p.sensitivity_threshold)
    return false;
  // uncertainty = min(u1, u2) (:319-321; linebase.cc:109-116; camera.cc:228-242)
  {
    double d1 = (c.zs + c.ze) / 2.0;
    double u1 = p.var2d * d1 / v1.fbar;
    double z2s = v2.P[8] * Xs.x + v2.P[9] * Xs.y + v2.P[10] * Xs.z + v2.P[11];
    double z2e = v2.P[8] * Xe.x + v2.P[9] * Xe.y + v2.P[10] * Xe.z + v2.P[11];
    double u2 = p.var2d * ((z2s + z2e) / 2.0) / v2.fbar;
    c.unc = smin(u1, u2);
  }
  // test_line_inside_ranges (functions.cc:8-26)
  if (p.ranges_flag) {
    if (Xs.x < p.rlo[0] || Xs.x > p.rhi[0] || Xs.y < p.rlo[1] || Xs.y > p.rhi[1] || Xs.z < p.rlo[2] || Xs.z > p.rhi[2])
      return false;
    if (Xe.x < p.rlo[0] || Xe.x > p.rhi[0] || Xe.y < p.rlo[1] || Xe.y > p.rhi[1] || Xe.z < p.rlo[2] || Xe.z > p.rhi[2])
      return false;
  }
  c.s = Xs;
  c.e = Xe;
  return true;
}

// Pair score of candidates (i, j) of one node (global_line_triangulator.cc:91-104):
// min(LineLinker3d::compute_score(l_i, l_j), LineLinker2d::compute_score(proj_{view j}(l_i), seg_j)),
// 0 when either is 0.
LM_D double pair_score(const TriParams &p, const seg<vec3<double>> &Li, vec3<double> di, double zsi, double zei,
                       const Slab &sl, int j, uint32_t vj) {
  // 3d: angle (line_linker.cc:185-192) then scale-invariant endpoint distance (:269-277)
  const LinkerDev<double> &c3 = p.l3d;
  double score3 = 1.0;
  {
    double cs = fabs(di.x * sl.dx[j] + di.y * sl.dy[j] + di.z * sl.dz[j]);
    double angle = acos(cs) * consts<double>::rad2deg();
    score3 = smin(score3, thresh0(expscore(angle, c3.th_angle * c3.mult), c3.score_th));
    if (score3 < c3.score_th) return 0.0;
    vec3<double> sj = mk3(sl.sx[j], sl.sy[j], sl.sz[j]), ej = mk3(sl.ex[j], sl.ey[j], sl.ez[j]);
    double ds = norm(Li.s - sj), de = norm(Li.e - ej);
    double dist = smax(ds / (zsi + consts<double>::eps()), de / (zei + consts<double>::eps()));
    score3 = smin(score3, thresh0(expscore(dist, c3.th_scaleinv * c3.mult), c3.score_th));
    if (score3 == 0.0) return 0.0;
  }
  // 2d: project l_i into the view of candidate j (linebase.cc:93-98) and score against its 2D segment
  const ViewD &v = p.views[vj];
  seg<vec2<double>> a, b;
  a.s = dehom(proj_h(v.P, Li.s));
  a.e = dehom(proj_h(v.P, Li.e));
  b.s = mk2(sl.q0[j], sl.q1[j]);
  b.e = mk2(sl.q2[j], sl.q3[j]);
  double score2 = linker_score<double, vec2<double>>(p.l2d, a, b, 1.0, false, 0.0, 0.0);
  if (score2 == 0.0) return 0.0;
  return smin(score3, score2);
}

__global__ void __launch_bounds__(kThreads) tri_node_kernel(const __grid_constant__ TriParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_wtot[kWarps];
  __shared__ int s_nvalid;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Slab sl;
  sl.carve(p.use_slab ? p.slab + (int64_t)blockIdx.x * p.slab_stride : reinterpret_cast<char *>(smem_raw), p.cap);

  for (int64_t node = p.node_begin + blockIdx.x; node < p.node_end; node += gridDim.x) {
    const uint32_t r0 = p.node_row_off[node], r1 = p.node_row_off[node + 1];
    const int nrows = (int)(r1 - r0);
    NodeRecord *rec = &p.nodes[node];
    if (nrows == 0) {
      if (tid < 9) rec->line[tid] = (tid == 8) ? -1.0 : 0.0;
      if (tid == 9) { rec->score = 0.0; rec->ng_view = 0; rec->ng_line = 0; rec->n_cand = 0; rec->n_valid = 0; }
      continue;
    }
    // ---------------- phase A: candidate generation with stable compaction -----------------
    const uint32_t v1i = p.node_view[node];
    const ViewD &v1 = p.views[v1i];
    Src src;
    src.l1 = ld_seg(&p.segs[node]);
    {
      double dx = src.l1.x - src.l1.z, dy = src.l1.y - src.l1.w;
      src.ok = !(sqrt(dx * dx + dy * dy) <= p.min_length_2d); // :166
      src.w1s = mat3_mul_h(v1.M, src.l1.x, src.l1.y);
      src.w1e = mat3_mul_h(v1.M, src.l1.z, src.l1.w);
      src.ray1s = normalized(src.w1s);
      src.ray1e = normalized(src.w1e);
      src.C1 = mk3(v1.C[0], v1.C[1], v1.C[2]);
    }
    int count = 0;
    for (int base = 0; base < nrows; base += kThreads) {
      const int r = base + tid;
      bool ok = false;
      Cand c;
      double4 l2;
      uint32_t ng = 0;
      if (r < nrows && src.ok) {
        ng = __ldg(&p.row_ng[r0 + r]);
        ok = gen_candidate(p, v1, src, ng >> 16, ng & 0xffffu, c, l2);
      }
      if (r < nrows && !ok) p.row_state[r0 + r] = 0;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_wtot[warp] = __popc(bal);
      __syncthreads();
      int woff = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        if (w < warp) woff += s_wtot[w];
        tot += s_wtot[w];
      }
      if (ok) {
        const int idx = count + woff + __popc(bal & ((1u << lane) - 1u));
        vec3<double> d = normalized(c.e - c.s);
        sl.sx[idx] = c.s.x; sl.sy[idx] = c.s.y; sl.sz[idx] = c.s.z;
        sl.ex[idx] = c.e.x; sl.ey[idx] = c.e.y; sl.ez[idx] = c.e.z;
        sl.dx[idx] = d.x; sl.dy[idx] = d.y; sl.dz[idx] = d.z;
        sl.zs[idx] = c.zs; sl.ze[idx] = c.ze; sl.unc[idx] = c.unc;
        sl.q0[idx] = l2.x; sl.q1[idx] = l2.y; sl.q2[idx] = l2.z; sl.q3[idx] = l2.w;
        sl.ng[idx] = ng;
        sl.row[idx] = (uint32_t)r;
      }
      count += tot;
      __syncthreads();
    }
    const int C = count;
    if (tid == 0) s_nvalid = 0;
    // ---------------- phase B: all-pairs scoring, one warp per candidate i ------------------
    for (int i = warp; i < C; i += kWarps) {
      seg<vec3<double>> Li;
      Li.s = mk3(sl.sx[i], sl.sy[i], sl.sz[i]);
      Li.e = mk3(sl.ex[i], sl.ey[i], sl.ez[i]);
      const vec3<double> di = mk3(sl.dx[i], sl.dy[i], sl.dz[i]);
      const double zsi = sl.zs[i], zei = sl.ze[i];
      const uint32_t vi = sl.ng[i] >> 16;
      double total = 0.0;
      uint32_t carry_view = 0xffffffffu;
      double carry_max = 0.0;
      for (int jb = 0; jb < C; jb += 32) {
        const int j = jb + lane;
        double s = 0.0;
        uint32_t vj = 0xfffffffeu;
        if (j < C) {
          vj = sl.ng[j] >> 16;
          if (j != i && vj != vi) s = pair_score(p, Li, di, zsi, zei, sl, j, vj);
        }
        if (vj == carry_view && carry_max > s) s = carry_max;
        // segmented inclusive max-scan over lanes with equal view (candidates are view-sorted)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, s, d);
          const uint32_t ov = __shfl_up_sync(0xffffffffu, vj, d);
          if (lane >= d && ov == vj && o > s) s = o;
        }
        uint32_t vnext = __shfl_down_sync(0xffffffffu, vj, 1);
        if (lane == 31) vnext = (j + 1 < C) ? (sl.ng[j + 1] >> 16) : 0xfffffffdu;
        const bool seg_end = (j < C) && (vnext != vj);
        // one image contributes its maximum once (:110-112); add in ascending-view order
        unsigned m = __ballot_sync(0xffffffffu, seg_end);
        while (m) {
          const int l = __ffs(m) - 1;
          total += __shfl_sync(0xffffffffu, s, l);
          m &= m - 1;
        }
        const uint32_t v31 = __shfl_sync(0xffffffffu, vj, 31);
        const double s31 = __shfl_sync(0xffffffffu, s, 31);
        const bool end31 = __shfl_sync(0xffffffffu, (int)seg_end, 31);
        if (!end31 && jb + 31 < C) { carry_view = v31; carry_max = s31; }
        else { carry_view = 0xffffffffu; carry_max = 0.0; }
      }
      if (lane == 0) sl.score[i] = total;
    }
    __syncthreads();
    // ---------------- phase C: valid connections + best candidate (:115-153) ----------------
    int nvalid_local = 0;
    for (int i = tid; i < C; i += kThreads) {
      const double sc = sl.score[i];
      bool valid = sc >= p.fullscore_th; // `if (score < fullscore_th) continue;`
      if (valid && C > p.max_valid_conns) {
        // rank in the (score, tri_id) descending order of std::greater<pair<double,int>> (:128-129)
        int rank = 0;
        for (int k = 0; k < C; ++k) {
          const double sk = sl.score[k];
          rank += (sk > sc) || (sk == sc && k > i);
        }
        valid = rank < p.max_valid_conns;
      }
      p.row_state[r0 + sl.row[i]] = valid ? 2 : 1;
      nvalid_local += valid;
      if (p.row_cand) {
        double *o = p.row_cand + (int64_t)(r0 + sl.row[i]) * 10;
        o[0] = sl.sx[i]; o[1] = sl.sy[i]; o[2] = sl.sz[i]; o[3] = sl.ex[i]; o[4] = sl.ey[i]; o[5] = sl.ez[i];
        o[6] = sl.zs[i]; o[7] = sl.ze[i]; o[8] = sl.unc[i]; o[9] = sc;
      }
    }
    if (nvalid_local) atomicAdd(&s_nvalid, nvalid_local);
    // best: first strict maximum from max_score = -1 (:145-153) == max score, lowest index on ties.
    if (warp == 0) {
      double bs = -1.0;
      int bi = -1;
      for (int i = lane; i < C; i += 32) {
        const double sc = sl.score[i];
        if (sc > bs) { bs = sc; bi = i; }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double os = __shfl_down_sync(0xffffffffu, bs, d);
        const int oi = __shfl_down_sync(0xffffffffu, bi, d);
        if (oi >= 0 && (os > bs || (os == bs && (bi < 0 || oi < bi)))) { bs = os; bi = oi; }
      }
      bi = __shfl_sync(0xffffffffu, bi, 0);
      if (bi >= 0) {
        if (lane < 3) rec->line[lane] = (lane == 0) ? sl.sx[bi] : (lane == 1 ? sl.sy[bi] : sl.sz[bi]);
        else if (lane < 6) rec->line[lane] = (lane == 3) ? sl.ex[bi] : (lane == 4 ? sl.ey[bi] : sl.ez[bi]);
        else if (lane == 6) rec->line[6] = sl.zs[bi];
        else if (lane == 7) rec->line[7] = sl.ze[bi];
        else if (lane == 8) rec->line[8] = sl.unc[bi];
        else if (lane == 9) rec->score = sl.score[bi];
        else if (lane == 10) { rec->ng_view = (int32_t)(sl.ng[bi] >> 16); rec->ng_line = (int32_t)(sl.ng[bi] & 0xffffu); }
      } else {
        if (lane < 9) rec->line[lane] = (lane == 8) ? -1.0 : 0.0;
        if (lane == 9) rec->score = 0.0;
        if (lane == 10) { rec->ng_view = 0; rec->ng_line = 0; }
      }
    }
    __syncthreads();
    if (tid == 0) {
      rec->n_cand = C;
      rec->n_valid = s_nvalid;
      if (C) atomicAdd(&p.counters[0], (unsigned long long)C);
      if (s_nvalid) atomicAdd(&p.counters[1], (unsigned long long)s_nvalid);
    }
    __syncthreads();
  }
}

void launch_tri_node_kernel(const TriParams &p, int grid, int block, size_t smem, cudaStream_t s) {
  (void)block;
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(tri_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  tri_node_kernel<<<grid, kThreads, smem, s>>>(p);
}

// ------------------------------------------------------------------------------------------------
// Match tables -> node-major rows. Flat row order = (source view asc, neighbour view asc, row), the
// order in which TriangulateImage appends to tris_ (base_line_triangulator.cc:74-100); a stable sort
// by node id then yields each node's candidates in reference order.
__global__ void expand_rows_kernel(const int32_t *__restrict__ pairs, const int64_t *__restrict__ blk_row_off,
                                   const int32_t *__restrict__ blk_src, const int32_t *__restrict__ blk_ng,
                                   const int64_t *__restrict__ blk_pair_off, int n_blocks,
                                   const int64_t *__restrict__ line_off, int64_t n_rows,
                                   uint32_t *__restrict__ key, uint32_t *__restrict__ val, int *err) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_blocks; // largest b with blk_row_off[b] <= r
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (blk_row_off[mid] <= r) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int64_t q = blk_pair_off[b] + (r - blk_row_off[b]);
    const int2 pr = reinterpret_cast<const int2 *>(pairs)[q];
    const int sv = blk_src[b], nv = blk_ng[b];
    const int64_t nl_src = line_off[sv + 1] - line_off[sv];
    const int64_t nl_ng = line_off[nv + 1] - line_off[nv];
    int line = pr.x, ngl = pr.y;
    if (line < 0 || line >= nl_src) { *err = 1; line = 0; }
    if (ngl < 0 || ngl >= nl_ng) { *err = 2; ngl = 0; }
    key[r] = (uint32_t)(line_off[sv] + line);
    val[r] = ((uint32_t)nv << 16) | (uint32_t)ngl;
  }
}
void launch_expand_rows(const int32_t *d_pairs, const int64_t *d_blk_row_off, const int32_t *d_blk_src_view,
                        const int32_t *d_blk_ng_view, const int64_t *d_blk_pair_off, int n_blocks,
                        const int64_t *d_line_off, int64_t n_rows, uint32_t *d_key, uint32_t *d_val, int *d_err,
                        cudaStream_t s) {
  if (n_rows == 0) return;
  int grid = (int)((n_rows + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  expand_rows_kernel<<<grid, 256, 0, s>>>(d_pairs, d_blk_row_off, d_blk_src_view, d_blk_ng_view, d_blk_pair_off,
                                          n_blocks, d_line_off, n_rows, d_key, d_val, d_err);
}

// TriangulateImageExhaustiveMatch (base_line_triangulator.cc:111-136): every line of the neighbour.
__global__ void expand_exhaustive_kernel(const int64_t *__restrict__ blk_row_off, const int32_t *__restrict__ blk_src,
                                         const int32_t *__restrict__ blk_ng, int n_blocks,
                                         const int64_t *__restrict__ line_off, int64_t n_rows,
                                         uint32_t *__restrict__ key, uint32_t *__restrict__ val) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_blocks;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (blk_row_off[mid] <= r) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int64_t q = r - blk_row_off[b];
    const int sv = blk_src[b], nv = blk_ng[b];
    const int64_t nl_ng = line_off[nv + 1] - line_off[nv];
    key[r] = (uint32_t)(line_off[sv] + q / nl_ng);
    val[r] = ((uint32_t)nv << 16) | (uint32_t)(q % nl_ng);
  }
}
void launch_expand_exhaustive(const int64_t *d_blk_row_off, const int32_t *d_blk_src_view,
                              const int32_t *d_blk_ng_view, int n_blocks, const int64_t *d_line_off, int64_t n_rows,
                              uint32_t *d_key, uint32_t *d_val, cudaStream_t s) {
  if (n_rows == 0) return;
  int grid = (int)((n_rows + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  expand_exhaustive_kernel<<<grid, 256, 0, s>>>(d_blk_row_off, d_blk_src_view, d_blk_ng_view, n_blocks, d_line_off,
                                                n_rows, d_key, d_val);
}

__global__ void node_offsets_kernel(const uint32_t *__restrict__ key, int64_t n_rows, int64_t n_nodes,
                                    uint32_t *__restrict__ off, unsigned int *max_rows) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n > n_nodes) return;
  // lower_bound(key, n)
  int64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (key[mid] < (uint32_t)n) lo = mid + 1; else hi = mid;
  }
  off[n] = (uint32_t)lo;
  if (n < n_nodes) {
    int64_t lo2 = lo, hi2 = n_rows;
    while (lo2 < hi2) {
      const int64_t mid = (lo2 + hi2) >> 1;
      if (key[mid] < (uint32_t)(n + 1)) lo2 = mid + 1; else hi2 = mid;
    }
    const unsigned int cnt = (unsigned int)(lo2 - lo);
    if (cnt) atomicMax(max_rows, cnt);
  }
}
void launch_node_offsets(const uint32_t *d_sorted_key, int64_t n_rows, int64_t n_nodes, uint32_t *d_node_row_off,
                         unsigned int *d_max_rows, cudaStream_t s) {
  const int grid = (int)((n_nodes + 1 + 255) / 256);
  node_offsets_kernel<<<grid, 256, 0, s>>>(d_sorted_key, n_rows, n_nodes, d_node_row_off, d_max_rows);
}

// valid_edges_ (global_line_triangulator.cc:130-142) in compact, node-major, candidate-ordered form.
__global__ void extract_nvalid_kernel(const NodeRecord *__restrict__ nodes, int64_t node_begin, int64_t n,
                                      uint32_t *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)nodes[node_begin + i].n_valid;
  if (i == n) out[i] = 0;
}
__global__ void compact_edges_kernel(const uint8_t *__restrict__ row_state, const uint32_t *__restrict__ row_ng,
                                     const uint32_t *__restrict__ node_row_off, const uint32_t *__restrict__ edge_off,
                                     int64_t node_begin, int64_t n, uint32_t *__restrict__ edge_ng) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (i >= n) return;
  const uint32_t r0 = node_row_off[node_begin + i], r1 = node_row_off[node_begin + i + 1];
  uint32_t base = edge_off[i];
  if (edge_off[i + 1] == base) return;
  for (uint32_t rb = r0; rb < r1; rb += 32) {
    const uint32_t r = rb + lane;
    const bool v = (r < r1) && row_state[r] == 2;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (v) edge_ng[base + __popc(m & ((1u << lane) - 1u))] = row_ng[r];
    base += __popc(m);
  }
}
// directed (src node, dst node) pairs of the compact edge list
__global__ void edge_pairs_kernel(const uint32_t *__restrict__ edge_off, const uint32_t *__restrict__ edge_ng,
                                  const int64_t *__restrict__ line_off, int64_t node_begin, int64_t n_nodes,
                                  int64_t n_edges, int64_t *__restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int64_t lo = 0, hi = n_nodes; // largest i with edge_off[i] <= e
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (edge_off[mid] <= (uint32_t)e) lo = mid; else hi = mid;
  }
  const uint32_t ng = edge_ng[e];
  out[2 * e] = node_begin + lo;
  out[2 * e + 1] = line_off[ng >> 16] + (ng & 0xffffu);
}
void launch_edge_pairs(const uint32_t *edge_off, const uint32_t *edge_ng, const int64_t *line_off,
                       int64_t node_begin, int64_t n_nodes, int64_t n_edges, int64_t *out, cudaStream_t s) {
  if (n_edges <= 0) return;
  edge_pairs_kernel<<<(int)((n_edges + 255) / 256), 256, 0, s>>>(edge_off, edge_ng, line_off, node_begin, n_nodes,
                                                                 n_edges, out);
}
void launch_extract_nvalid(const NodeRecord *nodes, int64_t node_begin, int64_t n, uint32_t *out, cudaStream_t s) {
  extract_nvalid_kernel<<<(int)((n + 1 + 255) / 256), 256, 0, s>>>(nodes, node_begin, n, out);
}
void launch_compact_edges_only(const uint8_t *row_state, const uint32_t *row_ng, const uint32_t *node_row_off,
                               const uint32_t *edge_off, int64_t node_begin, int64_t n, uint32_t *edge_ng,
                               cudaStream_t s) {
  if (n <= 0) return;
  compact_edges_kernel<<<(int)((n * 32 + 255) / 256), 256, 0, s>>>(row_state, row_ng, node_row_off, edge_off,
                                                                   node_begin, n, edge_ng);
}

// run_clustering edge weight (global_line_triangulator.cc:263-288): LineLinker3d::compute_score of the
// two best lines under set_to_spatial_merging().
__global__ void edge_weights_kernel(const __grid_constant__ EdgeParams p) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= p.n) return;
  const NodeRecord &a = p.nodes[p.edges[2 * e]];
  const NodeRecord &b = p.nodes[p.edges[2 * e + 1]];
  seg<vec3<double>> l1, l2;
  l1.s = mk3(a.line[0], a.line[1], a.line[2]); l1.e = mk3(a.line[3], a.line[4], a.line[5]);
  l2.s = mk3(b.line[0], b.line[1], b.line[2]); l2.e = mk3(b.line[3], b.line[4], b.line[5]);
  const double unc = smin(a.line[8], b.line[8]);
  p.weight[e] = linker_score<double, vec3<double>>(p.l3d, l1, l2, unc, true, a.line[6], a.line[7]);
}
void launch_edge_weights(const EdgeParams &p, cudaStream_t s) {
  if (p.n <= 0) return;
  const int grid = (int)((p.n + 127) / 128);
  edge_weights_kernel<<<grid, 128, 0, s>>>(p);
}

} // namespace lm
