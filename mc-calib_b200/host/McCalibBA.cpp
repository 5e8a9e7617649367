// McCalibBA.cpp — the reference's refinement stages on top of libmcba. Each refine* method walks the same
// containers in the same order as the reference builds its ceres::Problem (file:line in the comments), but
// appends to flat arrays instead of allocating one cost functor per corner, then calls mcba_create/mcba_solve
// (the replacement of ceres::Solve) and scatters the parameter blocks back into the std::array members.
#include "McCalibBA.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

#include "../../include/mcba.h"

namespace McCalib {

int g_cuda_device = 0;
int g_verbose = 1;
std::string g_last_error;
std::string g_problem_dump_dir;    // non-empty: every packed problem is written there before it is solved (mcba_problem_save)

// ---------------------------------------------------------------------------------------------- pose helpers
void rotVecToMat(const double* rv, double* R) {
  const double th2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
  if (th2 < 1e-30) { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
  const double th = std::sqrt(th2), c = std::cos(th), s = std::sin(th), oc = 1.0 - c;
  const double w[3] = {rv[0] / th, rv[1] / th, rv[2] / th};
  R[0] = c + oc * w[0] * w[0];         R[1] = -s * w[2] + oc * w[0] * w[1]; R[2] = s * w[1] + oc * w[0] * w[2];
  R[3] = s * w[2] + oc * w[1] * w[0];  R[4] = c + oc * w[1] * w[1];         R[5] = -s * w[0] + oc * w[1] * w[2];
  R[6] = -s * w[1] + oc * w[2] * w[0]; R[7] = s * w[0] + oc * w[2] * w[1];  R[8] = c + oc * w[2] * w[2];
}
void matToRotVec(const double* R, double* rv) {
  // through the unit quaternion (robust near pi)
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) { const double s = std::sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  if (q[0] < 0) for (double& v : q) v = -v;
  const double sn = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (sn < 1e-15) { rv[0] = rv[1] = rv[2] = 0.0; return; }
  const double ang = 2.0 * std::atan2(sn, q[0]);
  for (int i = 0; i < 3; ++i) rv[i] = q[1 + i] / sn * ang;
}
Pose6 composePose(const Pose6& o, const Pose6& in) {
  double Ro[9], Ri[9], R[9];
  rotVecToMat(o.data(), Ro); rotVecToMat(in.data(), Ri);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i * 3 + j] = Ro[i * 3] * Ri[j] + Ro[i * 3 + 1] * Ri[3 + j] + Ro[i * 3 + 2] * Ri[6 + j];
  Pose6 out{};
  matToRotVec(R, out.data());
  for (int i = 0; i < 3; ++i) out[3 + i] = Ro[i * 3] * in[3] + Ro[i * 3 + 1] * in[4] + Ro[i * 3 + 2] * in[5] + o[3 + i];
  return out;
}
Pose6 invertPose(const Pose6& p) {
  double R[9]; rotVecToMat(p.data(), R);
  Pose6 out{};
  for (int i = 0; i < 3; ++i) { out[i] = -p[i]; out[3 + i] = -(R[i] * p[3] + R[3 + i] * p[4] + R[6 + i] * p[5]); }
  return out;
}

// ---------------------------------------------------------------------------------------------- packing
namespace {
struct Packed {
  std::vector<float> u, v, pts_xyz;
  std::vector<int32_t> pt_idx, bobs_cam, bobs_pose, bobs_board, cam_model;
  std::vector<int64_t> bobs_ptr{0};
  std::vector<uint8_t> cam_is_ref, board_is_ref;
  std::vector<double> cam_pose, pose, board_pose, intrinsics;
  std::vector<double*> cam_pose_dst, pose_dst, board_pose_dst, intr_dst;   // where the blocks live in the McCalib objects
  std::map<int, int> cam_slot, board_slot, board_pt_base;
  int stage = 0;

  int camSlot(const std::shared_ptr<Camera>& cam, Pose6* rel_pose, bool is_ref) {
    auto it = cam_slot.find(cam->cam_idx_);
    if (it != cam_slot.end()) return it->second;
    const int s = (int)cam_model.size();
    cam_slot[cam->cam_idx_] = s;
    cam_model.push_back(cam->distortion_model_);
    cam_is_ref.push_back(is_ref ? 1 : 0);
    for (double x : cam->intrinsics_) intrinsics.push_back(x);
    intr_dst.push_back(cam->intrinsics_.data());
    for (int k = 0; k < 6; ++k) cam_pose.push_back(rel_pose ? (*rel_pose)[k] : 0.0);
    cam_pose_dst.push_back(rel_pose ? rel_pose->data() : nullptr);
    return s;
  }
  int boardSlot(const std::shared_ptr<Board>& board, Pose6* pose_in_object, bool is_ref) {
    auto it = board_slot.find(board->board_id_);
    if (it != board_slot.end()) return it->second;
    const int s = (int)board_is_ref.size();
    board_slot[board->board_id_] = s;
    board_is_ref.push_back(is_ref ? 1 : 0);
    for (int k = 0; k < 6; ++k) board_pose.push_back(pose_in_object ? (*pose_in_object)[k] : 0.0);
    board_pose_dst.push_back(pose_in_object ? pose_in_object->data() : nullptr);
    board_pt_base[board->board_id_] = (int)(pts_xyz.size() / 3);
    for (const Point3f& p : board->pts_3d_) { pts_xyz.push_back(p.x); pts_xyz.push_back(p.y); pts_xyz.push_back(p.z); }
    return s;
  }
  int poseSlot(Pose6* p) {
    const int s = (int)pose_dst.size();
    pose_dst.push_back(p->data());
    for (double x : *p) pose.push_back(x);
    return s;
  }
  void beginBobs(int cam, int pose_s, int board) { bobs_cam.push_back(cam); bobs_pose.push_back(pose_s); bobs_board.push_back(board); bobs_ptr.push_back((int64_t)u.size()); }
  void addObs(const Point2f& p, int pt) { u.push_back(p.x); v.push_back(p.y); pt_idx.push_back(pt); bobs_ptr.back() = (int64_t)u.size(); }

  // libmcba wants the board observations grouped by e-block: stable sort by pose slot (a frame with several objects
  // interleaves them, CameraGroup.cpp:383-386 walks object observations, not objects)
  void sortByPose() {
    const size_t nb = bobs_cam.size();
    bool sorted = true;
    for (size_t b = 1; b < nb; ++b) if (bobs_pose[b] < bobs_pose[b - 1]) { sorted = false; break; }
    if (sorted) return;
    std::vector<size_t> perm(nb);
    for (size_t b = 0; b < nb; ++b) perm[b] = b;
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return bobs_pose[a] < bobs_pose[b]; });
    Packed q;
    for (size_t k = 0; k < nb; ++k) {
      const size_t b = perm[k];
      q.beginBobs(bobs_cam[b], bobs_pose[b], bobs_board[b]);
      for (int64_t i = bobs_ptr[b]; i < bobs_ptr[b + 1]; ++i) { q.u.push_back(u[i]); q.v.push_back(v[i]); q.pt_idx.push_back(pt_idx[i]); }
      q.bobs_ptr.back() = (int64_t)q.u.size();
    }
    u.swap(q.u); v.swap(q.v); pt_idx.swap(q.pt_idx); bobs_ptr.swap(q.bobs_ptr);
    bobs_cam.swap(q.bobs_cam); bobs_pose.swap(q.bobs_pose); bobs_board.swap(q.bobs_board);
  }

  mcba_problem problem() const {
    mcba_problem P; std::memset(&P, 0, sizeof(P));
    P.stage = stage; P.n_cam = (int)cam_model.size(); P.n_pose = (int)pose_dst.size(); P.n_board = (int)board_is_ref.size();
    P.n_pts = (int)(pts_xyz.size() / 3); P.n_obs = (int64_t)u.size(); P.n_bobs = (int64_t)bobs_cam.size();
    P.u = u.data(); P.v = v.data(); P.pt_idx = pt_idx.data(); P.bobs_ptr = bobs_ptr.data();
    P.bobs_cam = bobs_cam.data(); P.bobs_pose = bobs_pose.data(); P.bobs_board = bobs_board.data();
    P.pts_xyz = pts_xyz.data(); P.cam_model = cam_model.data(); P.cam_is_ref = cam_is_ref.data();
    P.board_is_ref = board_is_ref.empty() ? nullptr : board_is_ref.data(); P.cam_problem = nullptr; P.n_problems = 1;
    return P;
  }
  mcba_params params() { return mcba_params{cam_pose.data(), pose.data(), board_pose.empty() ? nullptr : board_pose.data(), intrinsics.data()}; }
  void scatter(int st) {   // copy the refined blocks back into the McCalib objects (Ceres updates them in place)
    for (size_t i = 0; i < pose_dst.size(); ++i) std::copy_n(&pose[i * 6], 6, pose_dst[i]);
    if (st >= 3) for (size_t i = 0; i < cam_pose_dst.size(); ++i) if (cam_pose_dst[i]) std::copy_n(&cam_pose[i * 6], 6, cam_pose_dst[i]);
    if (st == 2 || st >= 4) for (size_t i = 0; i < board_pose_dst.size(); ++i) if (board_pose_dst[i]) std::copy_n(&board_pose[i * 6], 6, board_pose_dst[i]);
    if (st == 1 || st == 5) for (size_t i = 0; i < intr_dst.size(); ++i) std::copy_n(&intrinsics[i * 9], 9, intr_dst[i]);
  }
};

// ceres::Solve replacement with the reference's options (SPARSE_SCHUR, max_num_iterations, progress to stdout)
bool solveStages(Packed& pk, const std::vector<int>& stages, int nb_iterations) {
  if (pk.u.empty()) return true;   // nothing to refine (the reference would solve an empty problem)
  pk.stage = stages.front();
  pk.sortByPose();
  mcba_problem P = pk.problem();
  void* h = nullptr;
  int rc = mcba_create(&h, &P, g_cuda_device);
  for (size_t k = 0; rc == MCBA_OK && k < stages.size(); ++k) {
    if (k > 0) rc = mcba_set_stage(h, stages[k]);
    if (rc != MCBA_OK) break;
    mcba_options opt; mcba_default_options(&opt);
    opt.max_num_iterations = nb_iterations; opt.verbose = g_verbose;   // minimizer_progress_to_stdout = true in the reference
    mcba_params X = pk.params();
    mcba_summary summary;
    if (!g_problem_dump_dir.empty()) {   // checkpoint: the packed problem and the parameters this solve starts from
      static int n_dumped = 0;
      mcba_problem Pk = P; Pk.stage = stages[k];
      const std::string path = g_problem_dump_dir + "/problem_" + std::to_string(n_dumped++) + "_stage" + std::to_string(stages[k]) + ".mcba";
      if (mcba_problem_save(path.c_str(), &Pk, &X) != MCBA_OK) std::fprintf(stderr, "[McCalibBA] %s: %s\n", path.c_str(), mcba_io_last_error());
    }
    rc = mcba_solve(h, &X, &opt, &summary, nullptr, 0);
    if (rc == MCBA_OK) pk.scatter(stages[k]);
  }
  if (rc != MCBA_OK) { g_last_error = mcba_last_error(h); std::fprintf(stderr, "[McCalibBA] libmcba error %d: %s\n", rc, g_last_error.c_str()); }
  mcba_destroy(h);
  return rc == MCBA_OK;
}
}  // namespace

// ---------------------------------------------------------------------------------------------- stage 1
void Camera::refineIntrinsicCalibration(const int nb_iterations) {
  Packed pk;
  std::shared_ptr<Camera> self(std::shared_ptr<Camera>(), this);
  for (const auto& it : board_observations_) {                 // Camera.cpp:272
    auto bo = it.second.lock();
    if (!bo || !bo->valid_) continue;                          // :274
    auto board = bo->board_3d_.lock();
    if (!board) continue;
    const int cs = pk.camSlot(self, nullptr, false);
    const int bs = pk.boardSlot(board, nullptr, false);        // only for its point table
    pk.beginBobs(cs, pk.poseSlot(&bo->pose_), bs);             // blocks: board_obs->pose_, intrinsics_ (:288-290)
    for (size_t i = 0; i < bo->charuco_id_.size(); ++i) pk.addObs(bo->pts_2d_[i], pk.board_pt_base[board->board_id_] + bo->charuco_id_[i]);
  }
  pk.board_is_ref.clear(); pk.board_pose.clear(); pk.board_pose_dst.clear();   // stage 1 has no board block
  for (auto& b : pk.bobs_board) b = 0;
  solveStages(pk, {1}, nb_iterations);
}

// ---------------------------------------------------------------------------------------------- stage 2
void Object3D::refineObject(const int nb_iterations) {
  Packed pk;
  for (const auto& it_obj_obs : object_observations_) {        // Object3D.cpp:165
    auto oo = it_obj_obs.second.lock();
    if (!oo) continue;
    int pose_slot = -1;
    for (const auto& it_bo : oo->board_observations_) {        // :168
      auto bo = it_bo.second.lock();
      if (!bo || !bo->valid_) continue;                        // :170
      auto board = bo->board_3d_.lock(); auto cam = bo->cam_.lock();
      if (!board || !cam) continue;
      const int cs = pk.camSlot(cam, nullptr, false);          // intrinsics are constants here (:177-185)
      const int bs = pk.boardSlot(board, &relative_board_pose_[bo->board_id_], ref_board_id_ == bo->board_id_);   // :189-193
      if (pose_slot < 0) pose_slot = pk.poseSlot(&oo->pose_);  // blocks: object_obs->pose_, relative_board_pose_ (:206-209)
      pk.beginBobs(cs, pose_slot, bs);
      for (size_t i = 0; i < bo->charuco_id_.size(); ++i) pk.addObs(bo->pts_2d_[i], pk.board_pt_base[board->board_id_] + bo->charuco_id_[i]);
    }
  }
  solveStages(pk, {2}, nb_iterations);
  updateObjectPts();                                           // :225-242 regenerates pts_3d_
}

void Object3D::updateObjectPts() {
  for (const auto& it_board : boards_) {                       // Object3D.cpp:251
    auto board = it_board.second.lock();
    if (!board) continue;
    const Pose6& p = relative_board_pose_[board->board_id_];
    double R[9]; rotVecToMat(p.data(), R);
    for (size_t i = 0; i < board->pts_3d_.size(); ++i) {       // transform3DPts: float in, float out (geometrytools.cpp:346-352)
      const Point3f& q = board->pts_3d_[i];
      Point3f t;
      t.x = (float)(R[0] * q.x + R[1] * q.y + R[2] * q.z + p[3]);
      t.y = (float)(R[3] * q.x + R[4] * q.y + R[5] * q.z + p[4]);
      t.z = (float)(R[6] * q.x + R[7] * q.y + R[8] * q.z + p[5]);
      pts_3d_[pts_board_2_obj_[std::make_pair(board->board_id_, (int)i)]] = t;
    }
  }
}

void CameraGroupObs::updateObjObsPose() {
  for (const auto& it : object_observations_) {                // CameraGroupObs.cpp:209
    auto oo = it.second.lock();
    if (oo) oo->group_pose_ = object_pose_[oo->object_3d_id_]; // setPoseInGroupVec
  }
}

// ---------------------------------------------------------------------------------------------- stages 3-5
void CameraGroup::refineStage(int stage, int nb_iterations, bool then_intrinsics) {
  Packed pk;
  for (const auto& it_frame : frames_) {                                         // CameraGroup.cpp:371 (:196, :488)
    auto frame = it_frame.second.lock();
    if (!frame) continue;
    for (const auto& it_cgo : frame->cam_group_observations_) {                  // :375
      auto cgo = it_cgo.second.lock();
      if (!cgo || cam_group_idx_ != cgo->cam_group_idx_) continue;               // :378-381
      std::map<int, int> pose_slot_of_object;
      for (const auto& it_oo : cgo->object_observations_) {                      // :383-386
        auto oo = it_oo.second.lock();
        if (!oo) continue;
        auto object = oo->object_3d_.lock(); auto cam = oo->cam_.lock();
        if (!object || !cam) continue;
        const int cs = pk.camSlot(cam, &relative_camera_pose_[oo->camera_id_], id_ref_cam_ == cam->cam_idx_);   // :406-410
        auto ps = pose_slot_of_object.find(oo->object_3d_id_);
        if (ps == pose_slot_of_object.end())
          ps = pose_slot_of_object.emplace(oo->object_3d_id_, pk.poseSlot(&cgo->object_pose_[oo->object_3d_id_])).first;   // :446-448
        int cur_board = std::numeric_limits<int>::min();
        for (size_t i = 0; i < oo->pts_id_.size(); ++i) {                        // :411
          const std::pair<int, int>& bp = object->pts_obj_2_board_[oo->pts_id_[i]];   // :417-418
          if (stage == 3) {                                                      // object-frame points, no board block (:217-218, :238-239)
            if (i == 0) {
              if (pk.board_slot.find(-1 - object->obj_id_) == pk.board_slot.end()) {
                pk.board_slot[-1 - object->obj_id_] = 0;
                pk.board_pt_base[-1 - object->obj_id_] = (int)(pk.pts_xyz.size() / 3);
                for (const Point3f& p : object->pts_3d_) { pk.pts_xyz.push_back(p.x); pk.pts_xyz.push_back(p.y); pk.pts_xyz.push_back(p.z); }
              }
              pk.beginBobs(cs, ps->second, 0);
            }
            pk.addObs(oo->pts_2d_[i], pk.board_pt_base[-1 - object->obj_id_] + oo->pts_id_[i]);
            continue;
          }
          auto board = object->boards_[bp.first].lock();
          if (!board) continue;
          if (bp.first != cur_board) {                                           // a new run of corners of one board
            const int bs = pk.boardSlot(board, &object->relative_board_pose_[bp.first], object->ref_board_id_ == bp.first);   // :426-430
            pk.beginBobs(cs, ps->second, bs);
            cur_board = bp.first;
          }
          pk.addObs(oo->pts_2d_[i], pk.board_pt_base[bp.first] + bp.second);     // Board::pts_3d_, :423-425
        }
      }
    }
  }
  std::vector<int> stages{stage};
  if (then_intrinsics) stages.push_back(5);
  solveStages(pk, stages, nb_iterations);
}
void CameraGroup::refineCameraGroup(const int nb_iterations) { refineStage(3, nb_iterations, false); }
void CameraGroup::refineCameraGroupAndObjects(const int nb_iterations) { refineStage(4, nb_iterations, false); }
void CameraGroup::refineCameraGroupAndObjectsAndIntrinsics(const int nb_iterations) { refineStage(5, nb_iterations, false); }
void CameraGroup::refineCameraGroupAndObjectsThenIntrinsics(const int nb_iterations, bool with_intrinsics) { refineStage(4, nb_iterations, with_intrinsics); }

// ---------------------------------------------------------------------------------------------- Calibration
// stage-1 packing of every camera's valid board observations, camera-major (what the batched mode of libmcba wants:
// problem id == camera slot, one e-block per board observation)
namespace {
void packAllCamerasStage1(Calibration& calib, Packed& pk, std::vector<int32_t>* cam_problem) {
  for (const auto& it_cam : calib.cams_) {
    const std::shared_ptr<Camera>& cam = it_cam.second;
    for (const auto& it : cam->board_observations_) {                  // Camera.cpp:272
      auto bo = it.second.lock();
      if (!bo || !bo->valid_) continue;                                // :274
      auto board = bo->board_3d_.lock();
      if (!board || bo->charuco_id_.empty()) continue;
      const int cs = pk.camSlot(cam, nullptr, false);
      const int bs = pk.boardSlot(board, nullptr, false);              // only for its point table
      (void)bs;
      pk.beginBobs(cs, pk.poseSlot(&bo->pose_), 0);
      for (size_t i = 0; i < bo->charuco_id_.size(); ++i) pk.addObs(bo->pts_2d_[i], pk.board_pt_base[board->board_id_] + bo->charuco_id_[i]);
    }
  }
  pk.board_is_ref.clear(); pk.board_pose.clear(); pk.board_pose_dst.clear();   // stage 1 has no board block
  pk.stage = 1;
  if (cam_problem) { cam_problem->resize(pk.cam_model.size()); for (size_t c = 0; c < cam_problem->size(); ++c) (*cam_problem)[c] = (int32_t)c; }
}
}  // namespace

// McCalib.cpp:713-716 loops Camera::refineIntrinsicCalibration over the cameras: the per-camera problems are independent,
// so they are handed to libmcba as ONE batched launch (mcba_solve_batched: one Levenberg-Marquardt state per camera on
// the device); a single camera goes through the same path.
void Calibration::refineIntrinsicAndPoseAllCam() {
  g_cuda_device = cuda_device_;
  Packed pk; std::vector<int32_t> cam_problem;
  packAllCamerasStage1(*this, pk, &cam_problem);
  if (pk.u.empty()) return;
  mcba_problem P = pk.problem();
  P.cam_problem = cam_problem.data(); P.n_problems = (int32_t)cam_problem.size();
  void* h = nullptr;
  int rc = mcba_create(&h, &P, g_cuda_device);
  if (rc == MCBA_OK) {
    mcba_options opt; mcba_default_options(&opt);
    opt.max_num_iterations = nb_iterations_; opt.verbose = 0;          // one table per camera would interleave
    mcba_params X = pk.params();
    std::vector<mcba_summary> summaries(cam_problem.size());
    rc = mcba_solve_batched(h, &X, &opt, summaries.data());
    if (rc == MCBA_OK) pk.scatter(1);
  }
  if (rc != MCBA_OK) { g_last_error = h ? mcba_last_error(h) : "mcba_create failed"; std::fprintf(stderr, "[McCalibBA] libmcba error %d: %s\n", rc, g_last_error.c_str()); }
  mcba_destroy(h);
}

// McCalib.cpp:722-726 evaluates BoardObs::computeReprojectionError for every board observation (BoardObs.cpp:165-200:
// mean corner error under pose_ and the camera's intrinsics; it only logs). Returns the mean of those per-board means.
double Calibration::computeReproErrAllBoard() {
  g_cuda_device = cuda_device_;
  Packed pk;
  packAllCamerasStage1(*this, pk, nullptr);
  if (pk.u.empty()) return 0.0;
  mcba_problem P = pk.problem();
  void* h = nullptr;
  std::vector<double> err(pk.u.size());
  int rc = mcba_create(&h, &P, g_cuda_device);
  if (rc == MCBA_OK) { mcba_params X = pk.params(); rc = mcba_reprojection_errors(h, &X, err.data()); }
  if (rc != MCBA_OK) { g_last_error = h ? mcba_last_error(h) : "mcba_create failed"; mcba_destroy(h); return -1.0; }
  mcba_destroy(h);
  double sum = 0.0;
  const size_t nb = pk.bobs_cam.size();
  for (size_t b = 0; b < nb; ++b) {
    double s = 0.0;
    for (int64_t i = pk.bobs_ptr[b]; i < pk.bobs_ptr[b + 1]; ++i) s += err[i];
    sum += s / (double)(pk.bobs_ptr[b + 1] - pk.bobs_ptr[b]);
  }
  return sum / (double)nb;
}
void Calibration::refineAllObject3D() {
  g_cuda_device = cuda_device_;
  for (const auto& it : object_3d_) it.second->refineObject(nb_iterations_);            // McCalib.cpp:1015-1016
}
void Calibration::refineAllCameraGroup() {
  g_cuda_device = cuda_device_;
  for (const auto& it : cam_group_) it.second->refineCameraGroup(nb_iterations_);       // McCalib.cpp:1212-1213
  for (const auto& it : cams_group_obs_) it.second->updateObjObsPose();                 // :1216-1219
}
void Calibration::refineAllCameraGroupAndObjects() {
  g_cuda_device = cuda_device_;
  for (const auto& it : cam_group_) it.second->refineCameraGroupAndObjects(nb_iterations_);   // McCalib.cpp:1877-1878
  for (const auto& it : object_3d_) it.second->updateObjectPts();                             // :1881-1882
  for (const auto& it : cams_group_obs_) it.second->updateObjObsPose();                       // :1885-1886
}
void Calibration::refineAllCameraGroupAndObjectsAndIntrinsic() {
  g_cuda_device = cuda_device_;
  for (const auto& it : cam_group_) it.second->refineCameraGroupAndObjectsAndIntrinsics(nb_iterations_);   // McCalib.cpp:2349-2350
  for (const auto& it : object_3d_) it.second->updateObjectPts();
  for (const auto& it : cams_group_obs_) it.second->updateObjObsPose();
}

// Per-corner reprojection errors of every object observation of every camera group, through libmcba
// (mcba_reprojection_errors_f32 on the stage-3 packing: Object3D::pts_3d_ -> object pose -> camera pose -> projection)
// with the reference's own intermediate roundings (transform3DPts stores cv::Point3f, geometrytools.cpp:328-352;
// cv::projectPoints returns cv::Point2f; computeDistanceBetweenPoints stores floats, McCalib.cpp:2150-2160), so the
// files below hold the values the reference writes. One record per object observation, in the reference's walking order.
namespace {
struct ObsErrors { int group, frame, camera, object; std::vector<float> err; };
bool collectReprojectionErrors(Calibration& calib, std::vector<ObsErrors>& out) {
  for (const auto& it : calib.cam_group_) {
    CameraGroup& g = *it.second;
    Packed pk;
    std::vector<ObsErrors> recs;
    for (const auto& it_frame : g.frames_) {
      auto frame = it_frame.second.lock();
      if (!frame) continue;
      for (const auto& it_cgo : frame->cam_group_observations_) {
        auto cgo = it_cgo.second.lock();
        if (!cgo || g.cam_group_idx_ != cgo->cam_group_idx_) continue;
        std::map<int, int> pose_slot_of_object;
        for (const auto& it_oo : cgo->object_observations_) {
          auto oo = it_oo.second.lock();
          if (!oo) continue;
          auto object = oo->object_3d_.lock(); auto cam = oo->cam_.lock();
          if (!object || !cam || oo->pts_id_.empty()) continue;
          const int cs = pk.camSlot(cam, &g.relative_camera_pose_[oo->camera_id_], false);
          auto ps = pose_slot_of_object.find(oo->object_3d_id_);
          if (ps == pose_slot_of_object.end()) ps = pose_slot_of_object.emplace(oo->object_3d_id_, pk.poseSlot(&cgo->object_pose_[oo->object_3d_id_])).first;
          if (pk.board_pt_base.find(-1 - object->obj_id_) == pk.board_pt_base.end()) {
            pk.board_pt_base[-1 - object->obj_id_] = (int)(pk.pts_xyz.size() / 3);
            for (const Point3f& p : object->pts_3d_) { pk.pts_xyz.push_back(p.x); pk.pts_xyz.push_back(p.y); pk.pts_xyz.push_back(p.z); }
          }
          pk.beginBobs(cs, ps->second, (int)recs.size());   // the board slot is unused in stage 3: carry the record id through the sort
          for (size_t i = 0; i < oo->pts_id_.size(); ++i) pk.addObs(oo->pts_2d_[i], pk.board_pt_base[-1 - object->obj_id_] + oo->pts_id_[i]);
          recs.push_back(ObsErrors{g.cam_group_idx_, frame->frame_idx_, oo->camera_id_, oo->object_3d_id_, {}});
        }
      }
    }
    if (pk.u.empty()) continue;
    pk.stage = 3;
    pk.sortByPose();
    std::vector<int32_t> rec_of_bobs = pk.bobs_board;
    std::fill(pk.bobs_board.begin(), pk.bobs_board.end(), 0);
    mcba_problem P = pk.problem();
    void* h = nullptr;
    std::vector<float> err(pk.u.size());
    mcba_params X = pk.params();
    int rc = mcba_create(&h, &P, g_cuda_device);
    if (rc == MCBA_OK) rc = mcba_reprojection_errors_f32(h, &X, err.data());
    if (rc != MCBA_OK) { g_last_error = mcba_last_error(h); mcba_destroy(h); return false; }
    mcba_destroy(h);
    for (size_t b = 0; b + 1 < pk.bobs_ptr.size(); ++b)
      recs[rec_of_bobs[b]].err.assign(err.begin() + pk.bobs_ptr[b], err.begin() + pk.bobs_ptr[b + 1]);
    for (auto& r : recs) out.push_back(std::move(r));
  }
  return true;
}
}  // namespace

// Mean over object observations of the mean corner error (McCalib.cpp:2168-2245).
double Calibration::computeAvgReprojectionError() {
  g_cuda_device = cuda_device_;
  std::vector<ObsErrors> recs;
  if (!collectReprojectionErrors(*this, recs)) return std::numeric_limits<double>::quiet_NaN();
  double total = 0.0;
  for (const auto& r : recs) { double s = 0.0; for (float e : r.err) s += (double)e; total += s / (double)r.err.size(); }   // cv::mean
  return recs.empty() ? 0.0 : total / (double)recs.size();
}

// reprojection_error_data.yml (McCalib.cpp:2251-2341): camera_group_G { frame_F { camera_C { nb_pts, error_list },
// camera_list }, frame_list } with OpenCV-FileStorage matrices.
bool Calibration::saveReprojectionErrorToFile(const std::string& path) {
  g_cuda_device = cuda_device_;
  std::vector<ObsErrors> recs;
  if (!collectReprojectionErrors(*this, recs)) return false;
  std::ofstream f(path);
  if (!f) return false;
  f.precision(9);
  f << "%YAML:1.0\n---\nnb_camera_group: " << cam_group_.size() << "\n";
  auto mat = [&](const char* indent, const char* name, const char* dt, const std::vector<double>& v) {
    f << indent << name << ": !!opencv-matrix\n" << indent << "   rows: " << v.size() << "\n" << indent << "   cols: 1\n" << indent
      << "   dt: " << dt << "\n" << indent << "   data: [ ";
    for (size_t i = 0; i < v.size(); ++i) f << v[i] << (i + 1 < v.size() ? ", " : "");
    f << " ]\n";
  };
  auto matf = [&](const char* indent, const char* name, const std::vector<float>& v) {
    f << indent << name << ": !!opencv-matrix\n" << indent << "   rows: " << v.size() << "\n" << indent << "   cols: 1\n" << indent
      << "   dt: f\n" << indent << "   data: [ ";
    for (size_t i = 0; i < v.size(); ++i) f << v[i] << (i + 1 < v.size() ? ", " : "");
    f << " ]\n";
  };
  auto empty_mat = [&](const char* indent, const char* name) {   // what cv::FileStorage writes for an empty cv::Mat
    f << indent << name << ": !!opencv-matrix\n" << indent << "   rows: 0\n" << indent << "   cols: 0\n" << indent << "   dt: u\n" << indent << "   data: []\n";
  };
  std::vector<double> frame_list;   // the reference never clears it between groups (McCalib.cpp:2255): kept
  size_t k = 0;
  for (const auto& it : cam_group_) {
    const CameraGroup& g = *it.second;
    f << "camera_group_" << g.cam_group_idx_ << ":\n";
    for (const auto& it_frame : g.frames_) {                  // every frame of the group, with or without observations
      auto frame_ptr = it_frame.second.lock();
      if (!frame_ptr) continue;
      const int frame = frame_ptr->frame_idx_;
      f << "   frame_" << frame << ":\n";
      frame_list.push_back(frame);
      std::vector<double> camera_list;
      for (; k < recs.size() && recs[k].group == g.cam_group_idx_ && recs[k].frame == frame; ++k) {
        camera_list.push_back(recs[k].camera);
        f << "      camera_" << recs[k].camera << ":\n         nb_pts: " << recs[k].err.size() << "\n";
        matf("         ", "error_list", recs[k].err);
      }
      if (camera_list.empty()) empty_mat("      ", "camera_list"); else mat("      ", "camera_list", "i", camera_list);
    }
    if (frame_list.empty()) empty_mat("   ", "frame_list"); else mat("   ", "frame_list", "i", frame_list);
  }
  return true;
}

// CameraGroup::reproErrorCameraGroup (McCalib/src/CameraGroup.cpp:285-358), for every group
// (Calibration::reproErrorAllCamGroup, McCalib.cpp:1866-1869; called at apps/calibrate/src/calibrate.cpp:66): the mean
// error of each object observation - the reference only logs them (LOG_DEBUG); here they are returned as well.
// float accumulation of float errors, as the reference's `float sum_error`.
std::vector<Calibration::ObsMeanError> Calibration::reproErrorAllCamGroup() {
  g_cuda_device = cuda_device_;
  std::vector<ObsMeanError> out;
  std::vector<ObsErrors> recs;
  if (!collectReprojectionErrors(*this, recs)) return out;
  for (const auto& r : recs) {
    float sum_error = 0.f;
    for (float e : r.err) sum_error += e;
    out.push_back(ObsMeanError{r.group, r.frame, r.camera, r.object, sum_error / (float)r.err.size(), (int)r.err.size()});
  }
  return out;
}

// calibrated_objects_data.yml (McCalib.cpp:406-446): per object a 5 x N float matrix [X; Y; Z; board_id; pts_id].
bool Calibration::save3DObj(const std::string& path) const {
  std::ofstream f(path);
  if (!f) return false;
  f.precision(9);
  f << "%YAML:1.0\n---\n";
  for (const auto& it_obj : object_3d_) {
    const Object3D& obj = *it_obj.second;
    std::vector<std::array<float, 5>> cols;
    for (const auto& it_board : obj.boards_) {
      auto board = it_board.second.lock();
      if (!board) continue;
      for (size_t i = 0; i < board->pts_3d_.size(); ++i) {
        const Point3f& p = obj.pts_3d_[obj.pts_board_2_obj_.at(std::make_pair(board->board_id_, (int)i))];
        cols.push_back({p.x, p.y, p.z, (float)board->board_id_, (float)i});
      }
    }
    f << "object_" << obj.obj_id_ << ":\n   points: !!opencv-matrix\n      rows: 5\n      cols: " << cols.size() << "\n      dt: f\n      data: [ ";
    for (int r = 0; r < 5; ++r) for (size_t c = 0; c < cols.size(); ++c) f << cols[c][r] << ((r == 4 && c + 1 == cols.size()) ? "" : ", ");
    f << " ]\n";
  }
  return true;
}

// calibrated_objects_pose_data.yml (McCalib.cpp:452-483): per object a 6 x N double matrix of Object3DObs::pose_.
bool Calibration::save3DObjPose(const std::string& path) const {
  std::ofstream f(path);
  if (!f) return false;
  f.precision(17);
  f << "%YAML:1.0\n---\n";
  for (const auto& it_obj : object_3d_) {
    const Object3D& obj = *it_obj.second;
    std::vector<Pose6> poses;
    for (const auto& it_oo : obj.object_observations_) { auto oo = it_oo.second.lock(); if (oo) poses.push_back(oo->pose_); }
    f << "object_" << obj.obj_id_ << ":\n   poses: !!opencv-matrix\n      rows: 6\n      cols: " << poses.size() << "\n      dt: d\n      data: [ ";
    for (int r = 0; r < 6; ++r) for (size_t c = 0; c < poses.size(); ++c) f << poses[c][r] << ((r == 5 && c + 1 == poses.size()) ? "" : ", ");
    f << " ]\n";
  }
  return true;
}

// The handful of keys of the reference's OpenCV-FileStorage YAML that the refinement path reads (McCalib.cpp:39-79):
// number_iterations, fix_intrinsic, distortion_model, distortion_per_camera. "%YAML:1.0" files with "key: value"
// lines and "[a, b, ...]" lists; every other key is ignored.
bool Calibration::loadConfig(const std::string& yaml_path) {
  std::ifstream f(yaml_path);
  if (!f) return false;
  std::string line;
  while (std::getline(f, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    const size_t colon = line.find(':');
    if (colon == std::string::npos || line[0] == '%') continue;
    std::string key = line.substr(0, colon), val = line.substr(colon + 1);
    key.erase(0, key.find_first_not_of(" \t")); key.erase(key.find_last_not_of(" \t") + 1);
    for (char& ch : val) if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
    std::istringstream vs(val);
    if (key == "number_iterations") vs >> nb_iterations_;
    else if (key == "fix_intrinsic") vs >> fix_intrinsic_;
    else if (key == "ransac_threshold") vs >> ransac_thresh_;
    else if (key == "distortion_model") vs >> distortion_model_;
    else if (key == "distortion_per_camera") { distortion_per_camera_.clear(); int d; while (vs >> d) distortion_per_camera_.push_back(d); }
  }
  return true;
}

// calibrated_cameras_data.yml (McCalib.cpp:374-404): camera_matrix, distortion_vector, distortion_type, camera_group,
// img_width, img_height, camera_pose_matrix = inverse of relative_camera_pose_ (McCalib.cpp:398-399), as OpenCV FileStorage matrices.
bool Calibration::saveCamerasParams(const std::string& path) const {
  std::ofstream f(path);
  if (!f) return false;
  f.precision(17);
  f << "%YAML:1.0\n---\n";
  for (const auto& it_g : cam_group_) {
    f << "nb_camera: " << cams_.size() << "\n";
    for (const auto& it_c : it_g.second->cameras_) {
      auto cam = it_c.second.lock();
      if (!cam) continue;
      const auto& in = cam->intrinsics_;
      f << "camera_" << cam->cam_idx_ << ":\n";
      f << "   camera_matrix: !!opencv-matrix\n      rows: 3\n      cols: 3\n      dt: d\n      data: [ " << in[0] << ", 0., " << in[2] << ", 0., " << in[1] << ", " << in[3] << ", 0., 0., 1. ]\n";
      const int nd = cam->distortion_model_ == 0 ? 5 : 4;
      f << "   distortion_vector: !!opencv-matrix\n      rows: 1\n      cols: " << nd << "\n      dt: d\n      data: [ ";
      for (int k = 0; k < nd; ++k) f << in[4 + k] << (k + 1 < nd ? ", " : " ]\n");
      f << "   distortion_type: " << cam->distortion_model_ << "\n   camera_group: " << it_g.first << "\n";
      f << "   img_width: " << cam->im_cols_ << "\n   img_height: " << cam->im_rows_ << "\n";   // McCalib.cpp:395-396
      const Pose6 inv = invertPose(it_g.second->relative_camera_pose_.at(cam->cam_idx_));
      double R[9]; rotVecToMat(inv.data(), R);
      f << "   camera_pose_matrix: !!opencv-matrix\n      rows: 4\n      cols: 4\n      dt: d\n      data: [ ";
      for (int r = 0; r < 3; ++r) f << R[r * 3] << ", " << R[r * 3 + 1] << ", " << R[r * 3 + 2] << ", " << inv[3 + r] << ", ";
      f << "0., 0., 0., 1. ]\n";
    }
  }
  return true;
}

}  // namespace McCalib
