// TEST SCAFFOLDING (never shipped): gx_tree.cuh compiled as plain C++ so that the arithmetic of the
// general-tree step can be compared with the numpy oracle on the CPU before any GPU time is spent.
#include <vector>
#include "gx_tree.cuh"

extern "C" int tree_emu_model_size(void) { return (int)sizeof(GxTreeModel); }
extern "C" int tree_emu_scratch_floats(void) { return (int)GXT_SIZE; }

// q [n, nq], v [n, nv], warm [n, nv], ctrl [n, nu] row-major; updated in place
extern "C" int tree_emu_step(const GxTreeModel* m, int n, float* q, float* v, float* warm, const float* ctrl,
                             int forward_only, float* snap12, int* iters) {
  std::vector<float> sc(GXT_SIZE), sm(GXT_SM_FLOATS);
  for (int e = 0; e < n; ++e) {
    float* qe = q + (size_t)e * m->nq; float* ve = v + (size_t)e * m->nv; float* we = warm + (size_t)e * m->nv;
    GxTreeSnap sn;
    int it = 0;
    std::vector<float> qo(qe, qe + m->nq), vo(ve, ve + m->nv), wo(m->nv);
    gx_tree_step(*m, qe, ve, we, 1, ctrl + (size_t)e * m->nu, sc.data(), 1, sc.data(), sm.data(), 1, sm.data() + GXT_MOFF(m->nv), 1, forward_only != 0, qo.data(), vo.data(), wo.data(), sn, &it);
    for (int k = 0; k < m->nq; ++k) qe[k] = qo[k];
    for (int k = 0; k < m->nv; ++k) { ve[k] = vo[k]; we[k] = wo[k]; }
    if (snap12) { for (int k = 0; k < 3; ++k) snap12[12 * e + k] = sn.pos[k]; for (int k = 0; k < 9; ++k) snap12[12 * e + 3 + k] = sn.mat[k]; }
    if (iters) iters[e] = it;
  }
  return 0;
}
