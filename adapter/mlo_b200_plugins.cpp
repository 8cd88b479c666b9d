// adapter/mlo_b200_plugins.cpp  ->  libmlo_b200_plugins.so   (link: mp2p_icp, mrpt-maps, mrpt-poses, libmlo_b200.so)
//
// The MRPT-typed drop-in: registers `mlo_b200::ICP`, a subclass of mp2p_icp::ICP whose align()
// forwards to the C ABI of include/mlo_b200.h.  A user edits ONE line of the pipeline YAML
//     class_name: mlo_b200::ICP                     (pipelines/lidar3d-default.yaml:169)
// and runs   mola-lidar-odometry-cli ... -l libmlo_b200_plugins.so   (apps/mola-lidar-odometry-cli.cpp:93-95,553-561).
// Registration idiom: module/src/register.cpp:40-46.  Call site served: module/src/LidarOdometry.cpp:961-962.
//
// Built only where mp2p_icp and MRPT exist.  In this repository's environment they do not: CI checks the
// file with `g++ -fsyntax-only -I tests/mock_mp2p` against declaration-only headers (tests/test_adapter_syntax.py).
#if !defined(__has_include)
#error "this adapter needs __has_include (C++17)"
#elif __has_include(<mp2p_icp/ICP.h>)

#include <mp2p_icp/ICP.h>
#include <mp2p_icp/Matcher_Points_DistanceThreshold.h>
#include <mp2p_icp/Matcher_Point2Plane.h>
#include <mp2p_icp/Solver_GaussNewton.h>
#include <mp2p_icp/Solver_Horn.h>
#include <mrpt/core/exceptions.h>
#include <mrpt/core/initializer.h>
#include <mrpt/maps/CPointsMap.h>
#include <mrpt/poses/CPose3D.h>
#include <mrpt/rtti/CObject.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "mlo_b200.h"

namespace mlo_b200 {

// ---- MRPT <-> POD conversions ----------------------------------------------------------------------------------
static void to3x4(const mrpt::poses::CPose3D& p, double out[12]) {
  const auto& R = p.getRotationMatrix();
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) out[4 * r + c] = R(r, c);
  out[3] = p.x();
  out[7] = p.y();
  out[11] = p.z();
}
static mrpt::poses::CPose3D from3x4(const double in[12]) {
  mrpt::math::CMatrixDouble33 R;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R(r, c) = in[4 * r + c];
  const double t[3] = {in[3], in[7], in[11]};
  return mrpt::poses::CPose3D::FromRotationAndTranslation(R, t);
}
static void copy66(const mrpt::math::CMatrixDouble66& m, double out[36]) {
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) out[6 * r + c] = m(r, c);
}
static mrpt::math::CMatrixDouble66 from66(const double in[36]) {
  mrpt::math::CMatrixDouble66 m;
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) m(r, c) = in[6 * r + c];
  return m;
}

class ICP : public mp2p_icp::ICP {
  DEFINE_MRPT_OBJECT(ICP, mlo_b200)
 public:
  ICP() {
    if (mlo_create(0, &ctx_) != MLO_OK) THROW_EXCEPTION("mlo_b200: no sm_100 device (this library has no CPU path)");
  }
  ~ICP() override {
    if (map_) mlo_map_destroy(map_);
    if (ctx_) mlo_destroy(ctx_);
  }

  // Geometry of the device mirror of the global ("localmap") layer; upstream's defaults from
  // pipelines/lidar3d-default.yaml:228-242.  A pipeline with other values sets them through these members before the
  // first align (the metric_map_definition block is not visible from inside ICP).
  mlo_map_params map_params{MLO_MAP_HASHED_VOXEL_POINTS, 1.0f, 20u, 0.0f, 0.05f, 5u, 1u << 17};
  std::string global_layer = "localmap", local_layer = "decimated_for_icp";

  // same signature as the reference's call site, module/src/LidarOdometry.cpp:961-962
  void align(const mp2p_icp::metric_map_t& local, const mp2p_icp::metric_map_t& global, const mrpt::math::TPose3D& init,
             const mp2p_icp::Parameters& p, mp2p_icp::Results& out,
             const std::optional<mrpt::poses::CPose3DPDFGaussianInf>& prior = std::nullopt,
             mp2p_icp::LogRecord* outputDebugInfo = nullptr) override {
    (void)outputDebugInfo;
    // 1) parameters: the YAML blocks solvers / matchers / quality were parsed by the base class into matchers() and
    //    solvers(); their values are read, the objects are not run.
    mlo_icp_params q;
    mlo_icp_params_default(&q);
    q.max_iterations = p.maxIterations;
    q.min_abs_step_trans = p.minAbsStep_trans;
    q.min_abs_step_rot = p.minAbsStep_rot;
    mp2p_icp::Matcher_Points_DistanceThreshold* m_pt = nullptr;
    mp2p_icp::Matcher_Point2Plane* m_pl = nullptr;
    for (const auto& m : matchers()) {
      if (auto* a = dynamic_cast<mp2p_icp::Matcher_Points_DistanceThreshold*>(m.get())) m_pt = a;
      else if (auto* b = dynamic_cast<mp2p_icp::Matcher_Point2Plane*>(m.get())) m_pl = b;
      else THROW_EXCEPTION("mlo_b200::ICP: unsupported matcher class");
    }
    mp2p_icp::Solver_GaussNewton* gn = nullptr;
    bool horn = false;
    for (const auto& s : solvers()) {
      if (auto* a = dynamic_cast<mp2p_icp::Solver_GaussNewton*>(s.get())) gn = a;
      else if (dynamic_cast<mp2p_icp::Solver_Horn*>(s.get())) horn = true;
      else THROW_EXCEPTION("mlo_b200::ICP: unsupported solver class");
    }
    if (!gn && !horn) THROW_EXCEPTION("mlo_b200::ICP: no solver");
    q.matcher_mask = (m_pt ? MLO_MATCHER_PT2PT : 0u) | (m_pl ? MLO_MATCHER_PT2PL : 0u);
    q.solver = gn ? MLO_SOLVER_GAUSS_NEWTON : MLO_SOLVER_HORN;
    // runtime formulas (threshold, robustKernelParam: default.yaml:190,198) are functions of ICP_ITERATION: realise them
    // once per iteration index on the host and hand the tables over
    const size_t n_it = std::max<size_t>(1, std::min<size_t>(p.maxIterations, 300));
    std::vector<double> thr(n_it, 0.0), thr_pl(n_it, 0.0), kp(n_it, 1.0);
    for (size_t it = 0; it < n_it; it++) {
      if (auto* src = attachedSource()) {
        src->updateVariable("ICP_ITERATION", double(it));
        src->realize();
      }
      if (m_pt) thr[it] = m_pt->threshold;
      if (m_pl) thr_pl[it] = m_pl->distanceThreshold;
      if (gn) kp[it] = gn->robustKernelParam;
    }
    q.table_len = uint32_t(n_it);
    q.pt2pt_threshold_by_iter = thr.data();
    q.pt2pl_threshold_by_iter = thr_pl.data();
    q.kernel_param_by_iter = kp.data();
    if (m_pt) q.threshold_angular_deg = m_pt->thresholdAngularDeg;
    if (gn) {
      q.gn_max_iterations = gn->maxIterations;
      q.gn_min_delta = gn->minDelta;
      q.robust_kernel = int(gn->robustKernel);  // None / GemanMcClure / Cauchy share their numbering with mlo_robust_kernel
    }
    if (prior) {  // in.prior of LidarOdometry.cpp:859-877
      q.has_prior = 1;
      to3x4(prior->mean, q.prior_pose_3x4);
      copy66(prior->cov_inv, q.prior_info_6x6);
    }
    // 2) the iteration hook (LidarOdometry.cpp:923-952) is an arbitrary std::function and cannot run on the device: while
    //    one is installed the loop is cut at every iteration (max_iterations = 1 per device call) and the hook sees each
    //    intermediate solution exactly as upstream.  A caller that owns its hook passes it as DATA instead
    //    (mlo_icp_params.hook_*: thresholds + checkpoint, what the stock LidarOdometry hook computes) and keeps the whole
    //    loop on the device - that is what the MRPT-free host layer of this repository does (host/pipeline.hpp).
    // 3) global map: mirror the "localmap" layer on the device once, then feed only what the merge pipeline appended
    //    (the caller mutates it solely through that pipeline, LidarOdometry.cpp:1181-1200).
    syncDeviceMap(global);
    const auto pts = local.point_layer(local_layer);
    ASSERT_(pts);
    double init34[12];
    to3x4(mrpt::poses::CPose3D(init), init34);
    mlo_icp_result r;
    std::memset(&r, 0, sizeof(r));
    if (!iteration_hook_) {
      check(mlo_icp_align_soa(ctx_, pts->getPointsBufferRef_x().data(), pts->getPointsBufferRef_y().data(),
                              pts->getPointsBufferRef_z().data(), pts->size(), map_, init34, &q, &r));
    } else {
      // generic hook: one device iteration per call, the hook sees every intermediate solution exactly as upstream
      // (the stall test of each call compares with the previous solution only: upstream's extra comparison with the
      // solution before that - its 2-cycle guard - needs the whole loop in one call, i.e. the hook passed as data)
      uint32_t done = 0;
      double cur[12];
      std::memcpy(cur, init34, sizeof(cur));
      r.termination = MLO_TERM_MAX_ITERATIONS;
      while (done < p.maxIterations) {
        mlo_icp_params q1 = q;
        q1.max_iterations = 1;
        const size_t off = std::min<size_t>(done, n_it - 1);  // iteration `done` reads entry `done` of the tables
        q1.table_len = uint32_t(n_it - off);
        q1.pt2pt_threshold_by_iter = thr.data() + off;
        q1.pt2pl_threshold_by_iter = thr_pl.data() + off;
        q1.kernel_param_by_iter = kp.data() + off;
        check(mlo_icp_align_soa(ctx_, pts->getPointsBufferRef_x().data(), pts->getPointsBufferRef_y().data(),
                                pts->getPointsBufferRef_z().data(), pts->size(), map_, cur, &q1, &r));
        std::memcpy(cur, r.pose_3x4, sizeof(cur));
        if (r.termination != MLO_TERM_MAX_ITERATIONS) {  // NoPairings / SolverError / Stalled: the loop ended by itself
          done += r.n_iterations;
          break;
        }
        IterationHook_Input hi;
        IterationHook_Input::Solution sol;
        sol.optimalPose = from3x4(cur);
        hi.currentIteration = done;
        hi.currentSolution = &sol;
        if (iteration_hook_(hi).request_stop) {  // (upstream does not count the iteration that the hook stopped)
          r.termination = MLO_TERM_HOOK_REQUEST;
          break;
        }
        done += 1;
      }
      r.n_iterations = done;
    }
    out.optimal_tf.mean = from3x4(r.pose_3x4);
    double cov_ypr[36];
    mlo_cov_tangent_to_ypr(r.pose_3x4, r.cov_6x6, cov_ypr);  // Results::optimal_tf.cov is in the yaw/pitch/roll chart
    out.optimal_tf.cov = from66(cov_ypr);
    out.quality = r.quality;
    out.nIterations = r.n_iterations;
    out.terminationReason = static_cast<mp2p_icp::IterTermReason>(r.termination);
  }

 private:
  void check(int rc) const {
    // a negative status becomes the exception the reference's worker latches as fatal_error (LidarOdometry.cpp:614-619)
    if (rc != MLO_OK) THROW_EXCEPTION_FMT("mlo_b200: %s", mlo_last_error(ctx_));
  }
  // Device mirror of the global point layer.  The layer only ever grows by appends between aligns (merge pipeline) or
  // is replaced wholesale (map clear / load): append the new tail, or rebuild when it shrank.
  void syncDeviceMap(const mp2p_icp::metric_map_t& global) {
    const auto pts = global.point_layer(global_layer);
    ASSERT_(pts);
    if (!map_) check(mlo_map_create(ctx_, &map_params, &map_));
    const size_t n = pts->size();
    if (n < mirrored_ || pts.get() != mirrored_obj_) {
      check(mlo_map_clear(map_));
      mirrored_ = 0;
      mirrored_obj_ = pts.get();
    }
    if (n > mirrored_) {
      static const double I34[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // map points are already in the map frame
      check(mlo_map_insert_soa(map_, pts->getPointsBufferRef_x().data() + mirrored_, pts->getPointsBufferRef_y().data() + mirrored_,
                               pts->getPointsBufferRef_z().data() + mirrored_, n - mirrored_, I34));
      mirrored_ = n;
    }
  }
  mlo_ctx* ctx_ = nullptr;
  mlo_map* map_ = nullptr;
  size_t mirrored_ = 0;
  const void* mirrored_obj_ = nullptr;
};
IMPLEMENTS_MRPT_OBJECT(ICP, mp2p_icp::ICP, mlo_b200)

}  // namespace mlo_b200

// idiom of module/src/register.cpp:40-46
MRPT_INITIALIZER(register_mlo_b200_plugins) { mrpt::rtti::registerClass(CLASS_ID(mlo_b200::ICP)); }

#else
// mp2p_icp is not installed: nothing to build (the C ABI and the MRPT-free host layer do not need this file).
#endif
