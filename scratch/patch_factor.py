p='dune_ax1_b200/csrc/solver.cu'
s=open(p).read()
start=s.index('// block ILU0 of one sub-slab per CTA (bilu0_decomposition')
end=s.index('// ------------------------------------------------------------------------------------------\n// v = (LU)^-1 d for one sub-slab per CTA')
new=r'''// block ILU0 of one sub-slab per CTA (bilu0_decomposition: L_ij = A_ij * A_jj^-1, diagonal stored
// inverted). One quad per wavefront row. The 720 B Jacobian row of level L+2 is streamed into shared
// memory with cp.async while level L is eliminated; the finished [Dinv | U] rows of the last three
// levels -- all a row ever needs from its lower neighbours -- are kept in a shared-memory ring, so
// the critical path of a level contains no global load.
#define AX1_RING_ROW 82  // 80 doubles + 2 padding: 8 quads of a warp hit 8 different bank groups
__global__ void __launch_bounds__(256, 1) ilu_factor_kernel(Dev g, const double* __restrict__ A, double* LF, double* LB,
                                                            int slab_w, const int* __restrict__ lvl_tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, q = tid >> 2, r = tid & 3, nq = blockDim.x >> 2;
  double* stage = reinterpret_cast<double*>(smem_raw);            // [2][nq][AX1_ROW]
  double* ring = stage + 2 * (size_t)nq * AX1_ROW;                 // [3][nq][AX1_RING_ROW]
  const SlabGeom sg = slab_geom(g, slab_w, lvl_tab, reinterpret_cast<int*>(ring + 3 * (size_t)nq * AX1_RING_ROW));
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;
  const unsigned qmask = 0xFu << (tid & 28);

  auto issue = [&](int L) {
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const double* src = A + (size_t)w.v * AX1_ROW;
      double* dst = stage + ((size_t)(L & 1) * nq + q) * AX1_ROW;
      for (int c = r; c < AX1_ROW / 2; c += 4) cp_async16(dst + 2 * c, src + 2 * c);
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);
  for (int L = 0; L < nlev; ++L) {
    cp_async_wait<1>();
    __syncthreads();  // staged row of this level and the ring rows of level L-1 are visible
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    double row[9][4];
    if (w.v >= 0) {
      const int il = w.il, j = w.j;
      bool ex[9];
      const double* arow = stage + ((size_t)(L & 1) * nq + q) * AX1_ROW;
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        ex[s] = (il + di >= 0) && (il + di < wc) && (j + dj >= 0) && (j + dj < g.nvy);
        // expand the arrow block row (common.cuh), branch-free in the lane index
        const double* blk = arow + s * AX1_BLK;
        const double2 a = *reinterpret_cast<const double2*>(blk + (r == 3 ? 6 : 2 * r));
        const double2 b2 = *reinterpret_cast<const double2*>(blk + (r == 3 ? 8 : 2 * r));
        row[s][0] = (ex[s] && (r == 3 || r == 0)) ? a.x : 0.0;
        row[s][1] = ex[s] ? (r == 3 ? a.y : (r == 1 ? a.x : 0.0)) : 0.0;
        row[s][2] = ex[s] ? (r == 3 ? b2.x : (r == 2 ? a.x : 0.0)) : 0.0;
        row[s][3] = ex[s] ? (r == 3 ? b2.y : a.y) : 0.0;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (!ex[s]) continue;
        const int dis = s % 3 - 1, djs = s / 3 - 1;
        const int Ln = L + dis + 2 * djs;
        const int qn = (j + djs) - lev_jlo(Ln, wc);
        const double* kb = ring + ((size_t)(Ln % 3) * nq + qn) * AX1_RING_ROW;  // row k: [Dinv, U5..U8]
        double lrow[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double sum = 0.0;
#pragma unroll
          for (int m = 0; m < 4; ++m) sum += row[s][m] * kb[4 * m + c];
          lrow[c] = sum;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) row[s][c] = lrow[c];
#pragma unroll
        for (int tt = 5; tt < 9; ++tt) {
          const int di = dis + (tt % 3 - 1), dj = djs + (tt / 3 - 1);
          if (di < -1 || di > 1 || dj < -1 || dj > 1) continue;
          const int s2 = (dj + 1) * 3 + (di + 1);
          if (!ex[s2]) continue;
          const double* ub = kb + (tt - 4) * 16;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) sum += lrow[m] * ub[4 * m + c];
            row[s2][c] -= sum;
          }
        }
      }
      // invert the pivot block (every lane of the quad holds one row of it)
      double d[4][4], dinv[4][4];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) d[m][c] = __shfl_sync(qmask, row[4][c], (tid & 28) + m);
      invert4(d, dinv);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double val = dinv[0][c];
        if (r == 1) val = dinv[1][c];
        if (r == 2) val = dinv[2][c];
        if (r == 3) val = dinv[3][c];
        row[4][c] = val;
      }
      const size_t p = sg.row_off + (size_t)(sg.lvl[L] + q);
#pragma unroll
      for (int s = 0; s < 4; ++s) st4(LF + p * 64 + s * 16 + r * 4, row[s]);
#pragma unroll
      for (int s = 4; s < 9; ++s) st4(LB + p * 80 + (s - 4) * 16 + r * 4, row[s]);
    }
    __syncthreads();  // every read of ring slot L%3 (it held level L-3) is done
    if (w.v >= 0) {
      double* kr = ring + ((size_t)(L % 3) * nq + q) * AX1_RING_ROW;
#pragma unroll
      for (int s = 4; s < 9; ++s) st4(kr + (s - 4) * 16 + r * 4, row[s]);
    }
    issue(L + 2);
  }
  cp_async_wait<0>();
}

'''
s=s[:start]+new+s[end:]
helpers_start=s.index('__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {')
helpers_end=s.index('template <int D>\n__global__ void __launch_bounds__(256) ilu_apply_kernel')
helpers=s[helpers_start:helpers_end]
s=s[:helpers_start]+s[helpers_end:]
anchor=s.index('// block ILU0 of one sub-slab per CTA (bilu0_decomposition')
s=s[:anchor]+helpers+s[anchor:]
s=s.replace('''  const size_t smem = (size_t)(slab_w + 2 * g.nvy) * sizeof(int);
  ilu_factor_kernel<<<nslab, slab_threads(slab_w), smem, st>>>(g, A, LU, LU + (size_t)g.nv * 64, slab_w, lvl_tab);''','''  const int nt = slab_threads(slab_w), nq = nt / 4;
  const size_t smem = (size_t)nq * (2 * AX1_ROW + 3 * AX1_RING_ROW) * sizeof(double) + (size_t)(slab_w + 2 * g.nvy) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(ilu_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  ilu_factor_kernel<<<nslab, nt, smem, st>>>(g, A, LU, LU + (size_t)g.nv * 64, slab_w, lvl_tab);''')
open(p,'w').write(s)
