// Host-side harness around csrc/mcba_math.cuh (the arithmetic the CUDA kernels run), so that it can be
// checked against the oracle on a box without a GPU. TEST CODE: never linked into libmcba.
#include "../../mc-calib_b200/csrc/mcba_math.cuh"
using namespace mcba;

template <int STAGE> struct ArraySink {
  double* J; double* r;
  void put(int col, double ju, double jv) {
    typedef Cols<STAGE> C;
    if (col == C::R0) { r[0] = ju; r[1] = jv; return; }
    const int k = C::to_ref(col);
    J[k] = ju; J[C::K + k] = jv;
  }
};

template <int STAGE>
static double run(int model, int flags, const double* cam, const double* pose, const double* board,
                  const double* intr, const double* xyz, const double* uv, double huber_a, double* r, double* J,
                  double* cost_only) {
  typedef StageTraits<STAGE> T;
  double pc[PREP], po[PREP], pb[PREP], ctx[CTX_SIZE];
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  prep_block(cam ? cam : zero6, pc);
  prep_block(pose, po);
  prep_block(board ? board : zero6, pb);
  const bool act_c = T::CAM && !(flags & 1), act_b = T::BOARD && !(flags & 2);
  for (int e = 0; e < CTX_SIZE; ++e) if (e < CTX_NO || e >= CTX_IN) ctx[e] = ctx_entry_a(e, pc, act_c, po, pb, act_b, intr);
  for (int e = CTX_NO; e < CTX_IN; ++e) ctx[e] = ctx_entry_b(e, ctx, po, pb, act_b);
  ArraySink<STAGE> sink{J, r};
  const double hr = obs_jacobian<STAGE>(ctx, act_c, act_b, model, xyz, uv[0], uv[1], huber_a, sink);
  double ru, rv;
  *cost_only = obs_cost(ctx, model, xyz, uv[0], uv[1], huber_a, ru, rv);
  return hr;
}

extern "C" double mathtest_obs(int stage, int model, int flags, const double* cam, const double* pose,
                               const double* board, const double* intr, const double* xyz, const double* uv,
                               double huber_a, double* r, double* J, double* cost_only) {
  switch (stage) {
    case 1: return run<1>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 2: return run<2>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 3: return run<3>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 4: return run<4>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 5: return run<5>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
  }
  return -1;
}
extern "C" int mathtest_chol6(double* A, double* b) {
  if (!chol6(A)) return 0;
  chol6_fwd(A, b); chol6_bwd(A, b); return 1;
}
