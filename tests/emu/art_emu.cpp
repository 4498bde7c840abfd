// TEST SCAFFOLDING (never shipped, never linked into libguardx.so): compiles
// guardx_b200/csrc/gx_art.cuh as plain C++ with LANES = 1 so that the arithmetic of
// the articulated step can be compared with the numpy oracle on the CPU before any GPU
// time is spent.  The product path is the CUDA build of the same header.
#include <cstring>
#include "gx_art.cuh"

template <int NLEG, int K, int NB>
static int run(const GxArtModel* m, int n, float* q, float* v, float* warm, const float* ctrl, int forward_only,
               float* snap, int* iters) {
  const int nv = m->nv;
  for (int e = 0; e < n; ++e) {
    GxArtLane<1, NLEG, K, NB> st;
    float cr[3] = {0, 0, 0}, cl[NLEG][K];
    float* qe = q + (size_t)e * nv; float* ve = v + (size_t)e * nv; float* we = warm + (size_t)e * nv;
    for (int r = 0; r < 3; ++r) {
      st.qr[r] = qe[m->root_qadr + r]; st.vr[r] = ve[m->root_qadr + r]; st.wr[r] = we[m->root_qadr + r];
      if (m->root.dof[r].act >= 0) cr[r] = ctrl[(size_t)e * m->nu + m->root.dof[r].act];
    }
    for (int l = 0; l < NLEG; ++l)
      for (int i = 0; i < K; ++i) {
        const bool on = i < m->leg[l].nk;
        const int a = m->leg_qadr[l] + i;
        st.ql[l][i] = on ? qe[a] : 0.f; st.vl[l][i] = on ? ve[a] : 0.f; st.wl[l][i] = on ? we[a] : 0.f;
        cl[l][i] = (on && m->leg[l].dof[i].act >= 0) ? ctrl[(size_t)e * m->nu + m->leg[l].dof[i].act] : 0.f;
      }
    GxArtSnap sn;
    float qro[3], vro[3], qlo[NLEG][K], vlo[NLEG][K];
    int it = 0;
    gx_art_step<1, NLEG, K, NB>(*m, 0, 0xffffffffu, st, cr, cl, true, forward_only != 0, false, sn, qro, vro, qlo, vlo, &it);
    for (int r = 0; r < 3; ++r) { qe[m->root_qadr + r] = st.qr[r]; ve[m->root_qadr + r] = st.vr[r]; we[m->root_qadr + r] = st.wr[r]; }
    for (int l = 0; l < NLEG; ++l)
      for (int i = 0; i < m->leg[l].nk; ++i) {
        const int a = m->leg_qadr[l] + i;
        qe[a] = st.ql[l][i]; ve[a] = st.vl[l][i]; we[a] = st.wl[l][i];
      }
    if (snap) { snap[4 * e] = sn.px; snap[4 * e + 1] = sn.py; snap[4 * e + 2] = sn.r00; snap[4 * e + 3] = sn.r10; }
    if (iters) iters[e] = it;
  }
  return 0;
}

extern "C" int art_emu_model_size(void) { return (int)sizeof(GxArtModel); }

// q, v, warm: [n, nv] (robot dofs only, model order); ctrl: [n, nu]
extern "C" int art_emu_step(const GxArtModel* m, int nleg, int K, int NB, int n, float* q, float* v, float* warm,
                            const float* ctrl, int forward_only, float* snap, int* iters) {
  if (nleg == 4 && K == 2 && NB == 3) return run<4, 2, 3>(m, n, q, v, warm, ctrl, forward_only, snap, iters);
  if (nleg == 2 && K == 5 && NB == 3) return run<2, 5, 3>(m, n, q, v, warm, ctrl, forward_only, snap, iters);
  if (nleg == 1 && K == 2 && NB == 2) return run<1, 2, 2>(m, n, q, v, warm, ctrl, forward_only, snap, iters);
  return -1;
}
