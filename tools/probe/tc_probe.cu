// Stand-alone check + timing of score_tc_kernel (tcgen05 path) against a float64 host evaluation.
// Build: nvcc -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a \
//        -I diffco_b200/csrc -o tools/probe/tc_probe tools/probe/tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dc_score_tc.cuh"

namespace dc { long long g_launch_count = 0; }
using namespace dc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

static double urand(unsigned long long& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(s >> 11) / 9007199254740992.0; }
static double nrand(unsigned long long& s) { double u = urand(s) + 1e-12, v = urand(s); return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); }

static void fk7(const double* q, double* x) {
  double th = 0, px = 0, py = 0;
  for (int i = 0; i < 7; ++i) { th += q[i]; px += cos(th); py += sin(th); x[2 * i] = px; x[2 * i + 1] = py; }
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 2000, B = argc > 2 ? atoi(argv[2]) : 65536, D = 7, F = 14;
  const int near_frac_pct = argc > 3 ? atoi(argv[3]) : 0;  // % of queries placed next to a support
  const double gamma = 10.0;
  unsigned long long seed = argc > 5 ? strtoull(argv[5], nullptr, 10) : 1234;
  const double spread = argc > 6 ? atof(argv[6]) : 1.0;  // joint-angle range of supports and queries, in units of pi
  std::vector<double> Sq(N * D), Sx(N * F), W(N), Q((size_t)B * D);
  for (auto& v : Sq) v = (urand(seed) * 2 - 1) * M_PI * spread;
  for (int n = 0; n < N; ++n) fk7(&Sq[n * D], &Sx[n * F]);
  for (auto& v : W) v = nrand(seed);
  for (size_t b = 0; b < (size_t)B; ++b) {
    if ((int)(urand(seed) * 100) < near_frac_pct) {
      const int n = (int)(urand(seed) * N) % N;
      for (int i = 0; i < D; ++i) Q[b * D + i] = Sq[n * D + i] + 0.02 * nrand(seed);
    } else {
      for (int i = 0; i < D; ++i) Q[b * D + i] = (urand(seed) * 2 - 1) * M_PI * spread;
    }
  }
  if (near_frac_pct > 0) for (int i = 0; i < D; ++i) Q[5 * D + i] = Sq[17 * D + i];  // exact coincidence in fp32? (not exactly: FK in float)
  std::vector<float> sf(N * F), wf(N), qf((size_t)B * D), table(N * 16, 0.f);
  for (int i = 0; i < N * F; ++i) sf[i] = (float)Sx[i];
  for (int n = 0; n < N; ++n) { wf[n] = (float)W[n]; for (int f = 0; f < F; ++f) table[n * 16 + f] = -sf[n * F + f]; table[n * 16 + 14] = wf[n]; }
  for (size_t i = 0; i < qf.size(); ++i) qf[i] = (float)Q[i];

  const int nch = tc_n_chunks(N);
  float *d_s, *d_w, *d_q, *d_table, *d_out, *d_dbg; unsigned char* d_blob;
  CK(cudaMalloc(&d_s, sf.size() * 4)); CK(cudaMalloc(&d_w, wf.size() * 4)); CK(cudaMalloc(&d_q, qf.size() * 4));
  CK(cudaMalloc(&d_table, table.size() * 4)); CK(cudaMalloc(&d_blob, tc_blob_bytes(N)));
  CK(cudaMalloc(&d_out, (size_t)B * 8 * 4));
  const size_t dbg_floats = (size_t)128 * nch * TcLayout::NC + 128 * 32 + 128 * 16;
  CK(cudaMalloc(&d_dbg, dbg_floats * 4));
  CK(cudaMemcpy(d_s, sf.data(), sf.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_q, qf.data(), qf.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_blob, 0, tc_blob_bytes(N)));
  CK(cudaMemset(d_out, 0, (size_t)B * 8 * 4));
  if (launch_pack_supports_tc(d_s, d_w, N, F, (float)gamma, d_blob, 0) != 0) { printf("pack failed\n"); return 1; }
  CK(cudaDeviceSynchronize());

  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.fk.type = DC_FK_PLANAR_CHAIN; a.fk.dof = 7; a.fk.n_points = 7; a.fk.point_dim = 2; a.fk.n_links = 7;
  for (int i = 0; i < 7; ++i) a.fk.link_length[i] = 1.0;
  dc_kernel_desc kd{DC_K_RQ, 2, gamma};
  make_radial_consts<float>(kd, &a.rc);
  a.blob = d_blob; a.table = d_table; a.q = d_q; a.score = d_out; a.grad = d_out + 1; a.grad_out = nullptr; a.dbg = d_dbg;
  a.batch = B; a.score_ld = 8; a.grad_ld = 8; a.n_sv = N; a.n_feat = F; a.n_in = D; a.row_stride = 16; a.f_pad = 14;
  a.err_coef = argc > 4 ? (float)atof(argv[4]) : 1.0e-6f; a.tol_pair = 2e-7f;
  unsigned long long* d_stats; CK(cudaMalloc(&d_stats, 8)); CK(cudaMemset(d_stats, 0, 8)); a.stats = d_stats;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int st = launch_score_tc<TC_GRAD>(a, sms, 0);
  {
    int nb = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, score_tc_kernel<TC_GRAD>, TcLayout::THREADS, TcLayout::SM_BYTES);
    printf("occupancy: %d CTAs per SM (%d B dynamic smem per CTA)\n", nb, TcLayout::SM_BYTES);
  }
  printf("launch status %d, tiles %d chunks %d smem %d\n", st, a.n_tiles, a.n_chunks, TcLayout::SM_BYTES);
  {
    float tr[TcLayout::TRAILER_FLOATS];
    CK(cudaMemcpy(tr, tc_trailer(d_blob, N), sizeof(tr), cudaMemcpyDeviceToHost));
    printf("trailer: max|s|^2 %.3f  max|w s| %.3f  max|s_f| %.3f  Sa %g  tau*c0 %g  tau %g  inv_g %g  gamma %g  valid %g\n", tr[0], tr[1],
           tr[2], tr[3], tr[4], tr[5], tr[6], tr[7], tr[8]);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }

  std::vector<float> out((size_t)B * 8), dbg(dbg_floats);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(dbg.data(), d_dbg, dbg.size() * 4, cudaMemcpyDeviceToHost));

  // ---- rho of tile 0 vs exact (float inputs, double arithmetic) ------------------------------------------------
  {
    double max_abs = 0, max_rel_e = 0; int bad = 0;
    for (int i = 0; i < 128 && i < B; ++i) {
      double x[14];
      for (int f = 0; f < 14; ++f) x[f] = dbg[(size_t)128 * nch * TcLayout::NC + 128 * 32 + i * 16 + f];  // the device's own features
      double xx = 0; for (int f = 0; f < F; ++f) xx += x[f] * x[f];
      for (int n = 0; n < N; ++n) {
        double rho = 0, ss = 0; for (int f = 0; f < F; ++f) { double d = x[f] - sf[n * F + f]; rho += d * d; ss += (double)sf[n*F+f]*sf[n*F+f]; }
        const double got = dbg[(size_t)i * (nch * TcLayout::NC) + n];
        const double err = fabs(got - rho);
        if (err > max_abs) max_abs = err;
        if (err / (xx + ss) > max_rel_e) max_rel_e = err / (xx + ss);
        if (err > 1e-2 && bad < 5) { printf("  rho mismatch row %d sv %d: got %.6f want %.6f\n", i, n, got, rho); ++bad; }
      }
    }
    printf("GEMM1 rho (tile 0): max abs err %.3e, max err/(|x|^2+|s|^2) %.3e  (vs float64 on the device's float features)\n", max_abs, max_rel_e);
  }
  // ---- score / grad vs float64 ------------------------------------------------------------------------------
  const int NCHECK = B < 4096 ? B : 4096;
  double smax = 0, gmax = 0, serr = 0, gerr = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int b0 = pass == 0 ? 0 : (B - NCHECK);
    for (int b = b0; b < b0 + NCHECK; ++b) {
      double x[14], qd[7], gxv[14] = {0}, sc = 0; for (int k = 0; k < 7; ++k) qd[k] = qf[(size_t)b * 7 + k];
      fk7(qd, x);
      for (int n = 0; n < N; ++n) {
        double rho = 0, d[14]; for (int f = 0; f < F; ++f) { d[f] = x[f] - (double)sf[n * F + f]; rho += d[f] * d[f]; }
        const double u = 1.0 / (1.0 + gamma / 2 * rho), k = u * u;
        sc += wf[n] * k;
        const double c = -2 * gamma * wf[n] * k * u;
        for (int f = 0; f < F; ++f) gxv[f] += c * d[f];
      }
      double gq[7];
      for (int i = 0; i < 7; ++i) {
        double px = i ? x[2 * (i - 1)] : 0, py = i ? x[2 * (i - 1) + 1] : 0, acc = 0;
        for (int j = i; j < 7; ++j) acc += -gxv[2 * j] * (x[2 * j + 1] - py) + gxv[2 * j + 1] * (x[2 * j] - px);
        gq[i] = acc;
      }
      smax = fmax(smax, fabs(sc)); serr = fmax(serr, fabs(sc - out[(size_t)b * 8]));
      for (int i = 0; i < 7; ++i) { gmax = fmax(gmax, fabs(gq[i])); gerr = fmax(gerr, fabs(gq[i] - out[(size_t)b * 8 + 1 + i])); }
    }
  }
  printf("score: max|err| %.3e / max|ref| %.3e = %.3e    grad: %.3e / %.3e = %.3e   (gate 1e-5)\n", serr, smax, serr / smax, gerr, gmax, gerr / gmax);
  {
    unsigned long long near = 0; CK(cudaMemcpy(&near, d_stats, 8, cudaMemcpyDeviceToHost));
    printf("near pairs evaluated exactly: %llu of %.3e (%.4f %%), err_coef %.2e\n", near, (double)B * N, 100.0 * near / ((double)B * N), a.err_coef);
    a.stats = nullptr;
  }
  if (getenv("TC_QUICK")) return 0;

  if (getenv("TC_TRACE")) {
    long long* d_tr; CK(cudaMalloc(&d_tr, (2048 + 4 * 1024) * 8)); CK(cudaMemset(d_tr, 0, (2048 + 4 * 1024) * 8));
    a.dbg = nullptr; a.trace = d_tr;
    launch_score_tc<TC_GRAD>(a, sms, 0); CK(cudaDeviceSynchronize());
    launch_score_tc<TC_GRAD>(a, sms, 0); CK(cudaDeviceSynchronize());
    std::vector<long long> tr(64 * 16); CK(cudaMemcpy(tr.data(), d_tr, tr.size() * 8, cudaMemcpyDeviceToHost));
    const long long b0 = tr[8];
    printf("chunk | mma: ccwait g2iss g1iss(g+2) | query w0: start  b2full rho    ld     comp   stdone\n");
    for (int g = 0; g < 44; ++g) {
      printf("%3d  |", g);
      for (int k : {3, 4, 1}) printf(" %6lld", tr[g * 16 + k] ? tr[g * 16 + k] - b0 : -1);
      printf(" |");
      for (int k : {8, 9, 10, 11, 12, 13}) printf(" %6lld", tr[g * 16 + k] ? tr[g * 16 + k] - b0 : -1);
      printf("\n");
    }
    printf("tile events (cycles rel. to chunk-0 start of tile 0): lower half: loop start, loop end, drained, next FK done | owners: loop start, loop end, drained, G ready, records, stored\n");
    for (int ti = 0; ti < 3; ++ti) { printf("tile %d:", ti); for (int ev : {0, 1, 2, 3, 8, 9, 10, 11, 12, 13}) printf(" %7lld", tr[(50 + ev) * 16 + ti] ? tr[(50 + ev) * 16 + ti] - b0 : -1); printf("\n"); }
    printf("FK stage of each tile (lower half, warp 0): entered, configurations staged, float64 FK done, A operand published\n");
    for (int ti = 0; ti < 3; ++ti) { printf("tile %d:", ti); for (int ev : {7, 4, 5, 6}) printf(" %7lld", tr[(50 + ev) * 16 + ti] ? tr[(50 + ev) * 16 + ti] - b0 : -1); printf("\n"); }
    {
      std::vector<long long> cr(4 * 1024); CK(cudaMemcpy(cr.data(), d_tr + 2048, cr.size() * 8, cudaMemcpyDeviceToHost));
      const int grid = a.n_tiles < 2 * sms ? a.n_tiles : 2 * sms;
      long long tmin = -1, tmax = 0; int per_sm[256] = {0};
      for (int c = 0; c < grid; ++c) { if (tmin < 0 || cr[c * 4 + 1] < tmin) tmin = cr[c * 4 + 1]; if (cr[c * 4 + 2] > tmax) tmax = cr[c * 4 + 2]; per_sm[cr[c * 4] & 255]++; }
      int hist[8] = {0}; for (int i = 0; i < 256; ++i) if (per_sm[i]) hist[per_sm[i] < 7 ? per_sm[i] : 7]++;
      printf("CTA residency: grid %d, SMs with 1/2/3+ CTAs: %d/%d/%d, kernel span %.1f us\n", grid, hist[1], hist[2], hist[3] + hist[4], (tmax - tmin) * 1e-3);
      double d1 = 0, d2 = 0, s1 = 0, s2 = 0; int n1 = 0, n2 = 0; long long late = 0;
      for (int c = 0; c < grid; ++c) { const double st = (cr[c * 4 + 1] - tmin) * 1e-3, du = (cr[c * 4 + 2] - cr[c * 4 + 1]) * 1e-3; if (st > late) late = (long long)st;
        if (cr[c * 4 + 3] == 1) { d1 += du; s1 += st; ++n1; } else { d2 += du; s2 += st; ++n2; } }
      printf("  CTAs with 1 tile: %d, mean start %.1f us, mean duration %.1f us | with 2 tiles: %d, mean start %.1f us, mean duration %.1f us | latest start %lld us\n", n1, n1 ? s1 / n1 : 0., n1 ? d1 / n1 : 0., n2, n2 ? s2 / n2 : 0., n2 ? d2 / n2 : 0., late);
    }
    a.trace = nullptr;
  }
  // ---- timing ------------------------------------------------------------------------------------------------
  a.dbg = nullptr;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) launch_score_tc<TC_GRAD>(a, sms, 0);
  cudaEventRecord(e0);
  const int IT = 20;
  for (int i = 0; i < IT; ++i) launch_score_tc<TC_GRAD>(a, sms, 0);
  cudaEventRecord(e1);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("timing launches failed: %s\n", cudaGetErrorString(e)); return 3; }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  printf("time per launch %.1f us -> %.3e evals/s (B=%d, N=%d)\n", 1e3 * ms / IT, (double)B * IT / (ms * 1e-3), B, N);
  return 0;
}
