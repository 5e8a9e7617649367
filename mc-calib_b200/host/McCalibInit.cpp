// McCalibInit.cpp — what sits immediately before and around the refinement path in the reference, on top of libmcba:
//   * pose initialisation: BoardObs::estimatePose / Object3DObs::estimatePose and their
//     Calibration::estimatePoseAll* loops become ONE batched mcba_pnp_ransac launch over every observation;
//     CameraGroup::computeObjPoseInCameraGroup, CameraGroupObs::computeObjectsPose and getAverageRotation are small
//     host arithmetic and stay on the host;
//   * the on-disk formats either side of the path (rank 3): detected_keypoints_data.yml reader/writer and the
//     camera-parameter file that seeds the intrinsics, in cv::FileStorage's YAML dialect (OcvYaml.hpp).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../include/mcba_pnp.h"
#include "McCalibBA.hpp"
#include "OcvYaml.hpp"

namespace McCalib {

// ------------------------------------------------------------------------------------------ batched pose estimation
namespace {
struct PnpBatch {
  std::vector<float> u, v, pts_xyz;
  std::vector<int32_t> pt_idx, bobs_cam, cam_model;
  std::vector<int64_t> bobs_ptr{0};
  std::vector<double> intrinsics;
  std::map<int, int> cam_slot;
  std::map<const void*, int> table_base;   // point table (a Board or an Object3D) -> first row in pts_xyz
  int camSlot(const Camera& c) {
    auto it = cam_slot.find(c.cam_idx_);
    if (it != cam_slot.end()) return it->second;
    const int s = (int)cam_model.size();
    cam_slot[c.cam_idx_] = s;
    cam_model.push_back(c.distortion_model_);
    for (double x : c.intrinsics_) intrinsics.push_back(x);
    return s;
  }
  int tableBase(const void* key, const std::vector<Point3f>& pts) {
    auto it = table_base.find(key);
    if (it != table_base.end()) return it->second;
    const int b = (int)(pts_xyz.size() / 3);
    table_base[key] = b;
    for (const Point3f& p : pts) { pts_xyz.push_back(p.x); pts_xyz.push_back(p.y); pts_xyz.push_back(p.z); }
    return b;
  }
  void add(const Camera& c, int base, const std::vector<Point2f>& pts_2d, const std::vector<int>& ids) {
    bobs_cam.push_back(camSlot(c));
    for (size_t i = 0; i < ids.size(); ++i) { u.push_back(pts_2d[i].x); v.push_back(pts_2d[i].y); pt_idx.push_back(base + ids[i]); }
    bobs_ptr.push_back((int64_t)u.size());
  }
  // ransacP3PDistortion(..., thresh, it, distortion_type) with its defaults p = 0.99, refine = true (geometrytools.hpp:51-58)
  bool run(float thresh, int it, unsigned long long seed, std::vector<double>& pose, std::vector<int32_t>& n_inl, std::vector<uint8_t>& mask) {
    const size_t nb = bobs_cam.size();
    pose.assign(nb * 6, 0.0); n_inl.assign(nb, 0); mask.assign(u.size(), 0);
    if (nb == 0) return true;
    mcba_pnp_problem P; std::memset(&P, 0, sizeof(P));
    P.n_obs = (int64_t)u.size(); P.u = u.data(); P.v = v.data(); P.pt_idx = pt_idx.data();
    P.n_bobs = (int64_t)nb; P.bobs_ptr = bobs_ptr.data(); P.bobs_cam = bobs_cam.data();
    P.n_pts = (int32_t)(pts_xyz.size() / 3); P.pts_xyz = pts_xyz.data();
    P.n_cam = (int32_t)cam_model.size(); P.cam_model = cam_model.data(); P.intrinsics = intrinsics.data();
    mcba_pnp_options opt; mcba_pnp_default_options(&opt);
    opt.threshold_px = thresh; opt.max_iterations = it; opt.seed = seed;
    mcba_pnp_result R; std::memset(&R, 0, sizeof(R));
    R.pose = pose.data(); R.n_inliers = n_inl.data(); R.inlier_mask = mask.data();
    const int rc = mcba_pnp_ransac(&P, &opt, g_cuda_device, &R);
    if (rc != MCBA_OK) {
      g_last_error = mcba_pnp_last_error();
      std::fprintf(stderr, "[McCalibBA] mcba_pnp_ransac error %d: %s\n", rc, g_last_error.c_str());
      return false;
    }
    return true;
  }
};

// BoardObs.cpp:143-161: pose, validity, and only the inliers are kept
void applyBoardResult(BoardObs& bo, const double* pose, int n_inliers, const uint8_t* mask) {
  std::copy_n(pose, 6, bo.pose_.data());
  if (n_inliers < 4) bo.valid_ = false;
  std::vector<Point2f> new_pts; std::vector<int> new_ids;
  for (size_t i = 0; i < bo.charuco_id_.size(); ++i) if (mask[i]) { new_pts.push_back(bo.pts_2d_[i]); new_ids.push_back(bo.charuco_id_[i]); }
  bo.pts_2d_.swap(new_pts); bo.charuco_id_.swap(new_ids);
}
}  // namespace

void BoardObs::estimatePose(const float ransac_thresh, const int ransac_iterations) {
  auto cam = cam_.lock(); auto board = board_3d_.lock();
  if (!cam || !board) return;
  PnpBatch pb;
  pb.add(*cam, pb.tableBase(board.get(), board->pts_3d_), pts_2d_, charuco_id_);
  std::vector<double> pose; std::vector<int32_t> ni; std::vector<uint8_t> mask;
  if (pb.run(ransac_thresh, ransac_iterations, 0x9E37ULL * (unsigned long long)(frame_id_ + 1) + (unsigned long long)camera_id_ * 131 + (unsigned long long)board_id_, pose, ni, mask))
    applyBoardResult(*this, pose.data(), ni[0], mask.data());
}

void Object3DObs::estimatePose(const float ransac_thresh, const int ransac_iterations) {
  auto cam = cam_.lock(); auto obj = object_3d_.lock();
  if (!cam || !obj) return;
  PnpBatch pb;
  pb.add(*cam, pb.tableBase(obj.get(), obj->pts_3d_), pts_2d_, pts_id_);
  std::vector<double> pose; std::vector<int32_t> ni; std::vector<uint8_t> mask;
  if (pb.run(ransac_thresh, ransac_iterations, 0x51EDULL * (unsigned long long)(frame_id_ + 1) + (unsigned long long)camera_id_ * 131 + (unsigned long long)object_3d_id_, pose, ni, mask))
    std::copy_n(pose.data(), 6, pose_.data());                 // Object3DObs.cpp:246-248: only the pose is set
}

void Calibration::estimatePoseAllBoards() {
  g_cuda_device = cuda_device_;
  PnpBatch pb; std::vector<BoardObs*> order;
  for (const auto& it : board_observations_) {                 // McCalib.cpp:704
    BoardObs& bo = *it.second;
    auto cam = bo.cam_.lock(); auto board = bo.board_3d_.lock();
    if (!cam || !board) continue;
    pb.add(*cam, pb.tableBase(board.get(), board->pts_3d_), bo.pts_2d_, bo.charuco_id_);
    order.push_back(&bo);
  }
  std::vector<double> pose; std::vector<int32_t> ni; std::vector<uint8_t> mask;
  if (!pb.run(ransac_thresh_, nb_iterations_, ransac_seed_, pose, ni, mask)) return;
  for (size_t b = 0; b < order.size(); ++b) applyBoardResult(*order[b], &pose[6 * b], ni[b], &mask[pb.bobs_ptr[b]]);
}

void Calibration::estimatePoseAllObjects() {
  g_cuda_device = cuda_device_;
  PnpBatch pb; std::vector<Object3DObs*> order;
  for (const auto& it : object_observations_) {                // McCalib.cpp:991
    Object3DObs& oo = *it.second;
    auto cam = oo.cam_.lock(); auto obj = oo.object_3d_.lock();
    if (!cam || !obj) continue;
    pb.add(*cam, pb.tableBase(obj.get(), obj->pts_3d_), oo.pts_2d_, oo.pts_id_);
    order.push_back(&oo);
  }
  std::vector<double> pose; std::vector<int32_t> ni; std::vector<uint8_t> mask;
  if (!pb.run(ransac_thresh_, nb_iterations_, ransac_seed_ + 1, pose, ni, mask)) return;
  for (size_t b = 0; b < order.size(); ++b) std::copy_n(&pose[6 * b], 6, order[b]->pose_.data());
}

// ------------------------------------------------------------------------------------------ group poses (host)
void CameraGroup::computeObjPoseInCameraGroup() {
  for (const auto& it_frame : frames_) {                                        // CameraGroup.cpp:154
    auto frame = it_frame.second.lock();
    if (!frame) continue;
    for (const auto& it_cgo : frame->cam_group_observations_) {                 // :159-160
      auto cgo = it_cgo.second.lock();
      if (!cgo || cam_group_idx_ != cgo->cam_group_idx_) continue;              // :162-163
      for (const auto& it_oo : cgo->object_observations_) {                     // :167
        auto oo = it_oo.second.lock();
        if (!oo) continue;
        // pose_in_ref = pose_cam^-1 * pose_obj (:172-174), stored with setPoseInGroupMat
        oo->group_pose_ = composePose(invertPose(relative_camera_pose_[oo->camera_id_]), oo->pose_);
      }
    }
  }
}

void CameraGroupObs::insertObjectObservation(const std::shared_ptr<Object3DObs>& obs) {
  object_idx_.push_back(obs->object_3d_id_);
  object_observations_[(int)object_observations_.size()] = obs;
}

void CameraGroupObs::computeObjectsPose() {
  // groups of observations of the same object; std::unique only collapses ADJACENT duplicates (CameraGroupObs.cpp:45-48),
  // the map below merges the rest
  std::vector<int> uniq = object_idx_;
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  std::map<int, std::vector<int>> groups;
  for (int id : uniq) {
    std::vector<int>& g = groups[id];
    if (!g.empty()) continue;   // the reference pushes the same indices again for a repeated id; its result is unchanged apart from the average's weights
    for (size_t j = 0; j < object_idx_.size(); ++j) if (object_idx_[j] == id) g.push_back((int)j);
  }
  auto group = cam_group_.lock();
  for (const auto& it : groups) {
    bool flag_ref_cam = false;
    Pose6 pose{};
    for (int k : it.second) {                                                   // :64-74
      auto oo = object_observations_[k].lock();
      if (group && oo && group->id_ref_cam_ == oo->camera_id_) { flag_ref_cam = true; pose = oo->group_pose_; }
    }
    if (!flag_ref_cam) {                                                        // :78-94
      std::vector<double> r1, r2, r3; double t[3] = {0, 0, 0};
      for (int k : it.second) {
        auto oo = object_observations_[k].lock();
        if (!oo) continue;
        r1.push_back(oo->group_pose_[0]); r2.push_back(oo->group_pose_[1]); r3.push_back(oo->group_pose_[2]);
        for (int c = 0; c < 3; ++c) t[c] += oo->group_pose_[3 + c];
      }
      const std::array<double, 3> r = getAverageRotation(r1, r2, r3, quaternion_averaging_);
      for (int c = 0; c < 3; ++c) { pose[c] = r[c]; pose[3 + c] = t[c] / (double)it.second.size(); }
    }
    object_pose_[it.first] = pose;                                              // :97
    for (int k : it.second) { auto oo = object_observations_[k].lock(); if (oo) oo->group_pose_ = pose; }   // :99-103
  }
}

void Calibration::computeAllObjPoseInCameraGroup() {
  for (const auto& it : cam_group_) it.second->computeObjPoseInCameraGroup();   // McCalib.cpp:1636-1637
  for (const auto& it : cams_group_obs_) it.second->computeObjectsPose();       // :1639-1640
}

namespace {
double median(std::vector<double>& v) {                                         // geometrytools.cpp:697-703
  std::sort(v.begin(), v.end());
  const size_t mid = v.size() / 2;
  return v.size() % 2 == 0 ? (v[mid] + v[mid - 1]) / 2 : v[mid];
}
// largest eigenvector of a symmetric 4x4 (cyclic Jacobi); the reference takes column 0 of cv::SVD's U, i.e. the
// eigenvector of the largest eigenvalue of the positive semi-definite A, up to sign
void topEigenvector4(double A[4][4], double* q) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < 4; ++p)
      for (int r = p + 1; r < 4; ++r) {
        if (std::fabs(A[p][r]) < 1e-300) continue;
        const double theta = (A[r][r] - A[p][p]) / (2.0 * A[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) { const double a = A[k][p], b = A[k][r]; A[k][p] = c * a - s * b; A[k][r] = s * a + c * b; }
        for (int k = 0; k < 4; ++k) { const double a = A[p][k], b = A[r][k]; A[p][k] = c * a - s * b; A[r][k] = s * a + c * b; }
        for (int k = 0; k < 4; ++k) { const double a = V[k][p], b = V[k][r]; V[k][p] = c * a - s * b; V[k][r] = s * a + c * b; }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i) if (A[i][i] > A[best][best]) best = i;
  for (int k = 0; k < 4; ++k) q[k] = V[k][best];
}
}  // namespace

std::array<double, 4> convertRotationMatrixToQuaternion(const double* R) {
  std::array<double, 4> Q{};                                                    // x y z w
  const double trace = R[0] + R[4] + R[8];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    Q[3] = s * 0.5; s = 0.5 / s;
    Q[0] = (R[7] - R[5]) * s; Q[1] = (R[2] - R[6]) * s; Q[2] = (R[3] - R[1]) * s;
  } else {
    const int i = R[0] < R[4] ? (R[4] < R[8] ? 2 : 1) : (R[0] < R[8] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    Q[i] = s * 0.5; s = 0.5 / s;
    Q[3] = (R[k * 3 + j] - R[j * 3 + k]) * s;
    Q[j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
    Q[k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
  }
  return Q;
}

std::array<double, 9> convertQuaternionToRotationMatrix(const std::array<double, 4>& q) {
  const double q0 = q[3], q1 = q[0], q2 = q[1], q3 = q[2];                     // q0 = w
  return {2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2),
          2 * (q1 * q2 + q0 * q3), 2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1),
          2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 2 * (q0 * q0 + q3 * q3) - 1};
}

std::array<double, 3> getAverageRotation(std::vector<double>& r1, std::vector<double>& r2, std::vector<double>& r3,
                                         const bool use_quaternion_averaging) {
  std::array<double, 3> out{0, 0, 0};
  if (r1.empty()) return out;
  if (!use_quaternion_averaging) { out = {median(r1), median(r2), median(r3)}; return out; }   // :888-891
  double A[4][4] = {{0}};
  for (size_t n = 0; n < r1.size(); ++n) {
    const double rv[3] = {r1[n], r2[n], r3[n]};
    double R[9]; rotVecToMat(rv, R);
    std::array<double, 4> Q = convertRotationMatrixToQuaternion(R);
    if (Q[3] < 0) for (double& x : Q) x = -x;                                   // antipodal configurations (:861-864)
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) A[a][b] += Q[a] * Q[b];
  }
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) A[a][b] /= (double)r1.size();
  double q[4]; topEigenvector4(A, q);
  const std::array<double, 9> R = convertQuaternionToRotationMatrix({q[0], q[1], q[2], q[3]});   // the sign of q cancels
  matToRotVec(R.data(), out.data());
  return out;
}

// ------------------------------------------------------------------------------------------ on-disk formats
void Calibration::insertNewBoard(const int cam_idx, const int frame_idx, const int board_idx,
                                 const std::vector<Point2f>& pts_2d, const std::vector<int>& charuco_idx,
                                 const std::string& frame_path) {
  if (!cams_.count(cam_idx)) { auto c = std::make_shared<Camera>(); c->cam_idx_ = cam_idx; cams_[cam_idx] = c; }
  auto new_board = std::make_shared<BoardObs>();                               // McCalib.cpp:599-601
  new_board->camera_id_ = cam_idx; new_board->frame_id_ = frame_idx; new_board->board_id_ = board_idx;
  new_board->pts_2d_ = pts_2d; new_board->charuco_id_ = charuco_idx;
  new_board->cam_ = cams_[cam_idx];
  if (boards_3d_.count(board_idx)) new_board->board_3d_ = boards_3d_[board_idx];
  board_observations_[(int)board_observations_.size()] = new_board;            // :604
  auto& cam_obs = cams_[cam_idx]->board_observations_;                         // :607 Camera::insertNewBoard
  cam_obs[(int)cam_obs.size()] = new_board;
  auto it = frames_.find(frame_idx);                                           // :613-626
  if (it == frames_.end()) { auto f = std::make_shared<Frame>(); f->frame_idx_ = frame_idx; it = frames_.emplace(frame_idx, f).first; }
  auto& fobs = it->second->board_observations_;
  fobs[(int)fobs.size()] = new_board;
  it->second->frame_path_[cam_idx] = frame_path;
  // CameraObs (:628-641) groups the boards of one (camera, frame); nothing on the refinement path reads it
}

bool Calibration::loadDetectedKeypoints(const std::string& path) {
  ocvyaml::Node root; ocvyaml::Parser parser; std::string err;
  if (!parser.parseFile(path, root, &err)) { g_last_error = err; return false; }
  int nb_camera = root["nb_camera"].asInt((int)cams_.size());
  if ((int)cams_.size() > nb_camera) nb_camera = (int)cams_.size();
  for (int cam_idx = 0; cam_idx < nb_camera; ++cam_idx) {                      // McCalib.cpp:176
    const ocvyaml::Node& d = root["camera_" + std::to_string(cam_idx)];
    if (d.empty()) continue;
    const std::vector<int> frame_idxs = d["frame_idxs"].asInts(), board_idxs = d["board_idxs"].asInts();
    const std::vector<std::string> frame_paths = d["frame_paths"].asStrings();
    const ocvyaml::Node& pts = d["pts_2d"]; const ocvyaml::Node& ids = d["charuco_idxs"];
    if (frame_idxs.empty() || frame_idxs.size() != frame_paths.size() || frame_paths.size() != board_idxs.size() ||
        board_idxs.size() != pts.seq.size() || pts.seq.size() != ids.seq.size()) {   // the assert of :201-204
      g_last_error = "inconsistent keypoint arrays for camera_" + std::to_string(cam_idx);
      return false;
    }
    if (!cams_.count(cam_idx)) { auto c = std::make_shared<Camera>(); c->cam_idx_ = cam_idx; cams_[cam_idx] = c; }
    cams_[cam_idx]->im_cols_ = d["img_width"].asInt(); cams_[cam_idx]->im_rows_ = d["img_height"].asInt();   // :198-199
    for (size_t i = 0; i < frame_idxs.size(); ++i) {                           // :206-209
      const std::vector<double> xy = pts.seq[i].asDoubles();                   // vector<Point2f> = flat [x0, y0, x1, y1, ...]
      std::vector<Point2f> p(xy.size() / 2);
      for (size_t k = 0; k < p.size(); ++k) { p[k].x = (float)xy[2 * k]; p[k].y = (float)xy[2 * k + 1]; }
      insertNewBoard(cam_idx, frame_idxs[i], board_idxs[i], p, ids.seq[i].asInts(), frame_paths[i]);
    }
  }
  return true;
}

namespace {
void writeFlow(std::ofstream& f, const std::string& indent, const std::string& head, const std::vector<std::string>& items) {
  // cv::FileStorage wraps flow sequences at ~80 columns, continuation lines indented three more columns
  std::string line = indent + head + "[ ";
  for (size_t i = 0; i < items.size(); ++i) {
    const std::string piece = items[i] + (i + 1 < items.size() ? "," : "");
    if (line.size() + piece.size() + 1 > 78 && line.size() > indent.size() + 3) { f << line << "\n"; line = indent + "   "; }
    line += piece + " ";
  }
  f << line << "]\n";
}
std::string fmtFloat(float x) {   // shortest text that reads back to the same float, with the "." OpenCV puts on integers
  char buf[32];
  for (int prec = 1; prec <= 9; ++prec) { std::snprintf(buf, sizeof(buf), "%.*g", prec, (double)x); if ((float)std::strtod(buf, nullptr) == x) break; }
  std::string s(buf);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".";
  return s;
}
}  // namespace

bool Calibration::saveDetectedKeypoints(const std::string& path) const {
  std::ofstream f(path);
  if (!f) return false;
  f << "%YAML:1.0\n---\nnb_camera: " << cams_.size() << "\n";                  // McCalib.cpp:498
  for (const auto& it_cam : cams_) {
    const int cam_idx = it_cam.second->cam_idx_;
    f << "camera_" << cam_idx << ":\n   img_width: " << it_cam.second->im_cols_ << "\n   img_height: " << it_cam.second->im_rows_ << "\n";
    std::vector<std::string> frame_idxs, board_idxs, frame_paths;
    std::vector<const BoardObs*> obs;
    for (const auto& it_frame : frames_)                                       // :514-539 (frame, board, observation order)
      for (const auto& it_board : boards_3d_)
        for (const auto& it_bo : board_observations_) {
          const BoardObs& bo = *it_bo.second;
          if (bo.camera_id_ != cam_idx || bo.frame_id_ != it_frame.second->frame_idx_ || bo.board_id_ != it_board.first) continue;
          frame_idxs.push_back(std::to_string(bo.frame_id_)); board_idxs.push_back(std::to_string(bo.board_id_));
          auto fp = it_frame.second->frame_path_.find(cam_idx);
          frame_paths.push_back(fp == it_frame.second->frame_path_.end() ? std::string() : fp->second);
          obs.push_back(&bo);
        }
    writeFlow(f, "   ", "frame_idxs: ", frame_idxs);
    f << "   frame_paths:\n";
    for (const std::string& p : frame_paths) {   // quoted, with the two escapes cv::FileStorage itself uses
      f << "      - \"";
      for (char ch : p) { if (ch == '"' || ch == '\\') f << '\\'; f << ch; }
      f << "\"\n";
    }
    writeFlow(f, "   ", "board_idxs: ", board_idxs);
    f << "   pts_2d:\n";
    for (const BoardObs* bo : obs) {
      std::vector<std::string> it;
      for (const Point2f& p : bo->pts_2d_) { it.push_back(fmtFloat(p.x)); it.push_back(fmtFloat(p.y)); }
      writeFlow(f, "      ", "- ", it);
    }
    f << "   charuco_idxs:\n";
    for (const BoardObs* bo : obs) {
      std::vector<std::string> it;
      for (int id : bo->charuco_id_) it.push_back(std::to_string(id));
      writeFlow(f, "      ", "- ", it);
    }
  }
  return (bool)f;
}

bool Calibration::loadInitialIntrinsics(const std::string& path) {
  if (path.size() < 4 || path.substr(path.size() - 4) != ".yml") { g_last_error = "camera parameter file must be a .yml"; return false; }   // :667-669
  ocvyaml::Node root; ocvyaml::Parser parser; std::string err;
  if (!parser.parseFile(path, root, &err)) { g_last_error = err; return false; }
  for (const auto& it : cams_) {                                               // :678-687
    const ocvyaml::Node& n = root["camera_" + std::to_string(it.first)];
    int r = 0, c = 0; std::vector<double> K, D;
    if (!n["camera_matrix"].asMatrix(r, c, K) || r != 3 || c != 3) { g_last_error = "camera_matrix missing for camera_" + std::to_string(it.first); return false; }
    if (!n["distortion_vector"].asMatrix(r, c, D)) { g_last_error = "distortion_vector missing for camera_" + std::to_string(it.first); return false; }
    auto& in = it.second->intrinsics_;                                         // Camera::setIntrinsics: fx fy u0 v0 | distortion
    in = {K[0], K[4], K[2], K[5], 0, 0, 0, 0, 0};
    for (size_t k = 0; k < D.size() && k < 5; ++k) in[4 + k] = D[k];
    if (!n["distortion_type"].empty()) it.second->distortion_model_ = n["distortion_type"].asInt();
  }
  return true;
}

}  // namespace McCalib

// ------------------------------------------------------------------------------------------ workflow head
namespace McCalib {

bool Calibration::initialization(const std::string& config_path) {
  if (config_path.size() < 4 || config_path.substr(config_path.size() - 4) != ".yml") { g_last_error = "config must be a .yml"; return false; }   // :31-34
  ocvyaml::Node fs; ocvyaml::Parser parser; std::string err;
  if (!parser.parseFile(config_path, fs, &err)) { g_last_error = err; return false; }
  const int nb_camera = fs["number_camera"].asInt(0);                          // :41-44
  int nb_board = fs["number_board"].asInt(0);                                  // :46-49
  if (nb_camera <= 0 || nb_board <= 0) { g_last_error = "number_camera and number_board must be positive"; return false; }
  const int nb_x_square = fs["number_x_square"].asInt(0), nb_y_square = fs["number_y_square"].asInt(0);
  // quaternion_averaging: the reference asks for the key "quaternion_averaging:" (with the colon, McCalib.cpp:57), which
  // no file has, so its quaternion_averaging_ always keeps the default (1); same here
  if (!fs["ransac_threshold"].empty()) ransac_thresh_ = (float)fs["ransac_threshold"].asDouble(10.0);
  if (!fs["number_iterations"].empty()) nb_iterations_ = fs["number_iterations"].asInt(1000);
  distortion_model_ = fs["distortion_model"].asInt(0);
  distortion_per_camera_ = fs["distortion_per_camera"].asInts();
  std::vector<int> boards_index = fs["boards_index"].asInts();
  cam_params_path_ = fs["cam_params_path"].scalar; keypoints_path_ = fs["keypoints_path"].scalar; save_path_ = fs["save_path"].scalar;
  std::vector<double> square_size_per_board = fs["square_size_per_board"].asDoubles();
  std::vector<int> nx_per_board = fs["number_x_square_per_board"].asInts(), ny_per_board = fs["number_y_square_per_board"].asInts();
  fix_intrinsic_ = fs["fix_intrinsic"].asInt(0);
  if (!boards_index.empty()) nb_board = (int)boards_index.size();              // :82-84
  int max_board_idx = nb_board - 1;
  if (!boards_index.empty()) max_board_idx = *std::max_element(boards_index.begin(), boards_index.end());
  if (nx_per_board.empty()) { nx_per_board.assign(max_board_idx + 1, nb_x_square); ny_per_board.assign(max_board_idx + 1, nb_y_square); }   // :89-92
  if (distortion_per_camera_.empty()) distortion_per_camera_.assign(nb_camera, distortion_model_);   // :105-106
  if ((int)distortion_per_camera_.size() < nb_camera) { g_last_error = "distortion_per_camera shorter than number_camera"; return false; }
  for (int i = 0; i < nb_camera; ++i) {                                        // :109-113
    auto cam = std::make_shared<Camera>();
    cam->cam_idx_ = i; cam->distortion_model_ = distortion_per_camera_[i];
    cams_[i] = cam;
  }
  if (boards_index.empty()) { boards_index.resize(nb_board); for (int b = 0; b < nb_board; ++b) boards_index[b] = b; }   // :116-119
  for (int board_idx = 0; board_idx < nb_board; ++board_idx) {                 // :125-134 + Board.cpp:22-72
    int nx, ny; double sq;
    if (square_size_per_board.empty()) { nx = nb_x_square; ny = nb_y_square; sq = fs["square_size"].asDouble(0.0); }
    else {
      const int bi = boards_index[board_idx];
      if (bi >= (int)nx_per_board.size() || bi >= (int)ny_per_board.size() || bi >= (int)square_size_per_board.size()) { g_last_error = "per-board lists shorter than the board index"; return false; }
      nx = nx_per_board[bi]; ny = ny_per_board[bi]; sq = square_size_per_board[bi];
    }
    auto board = std::make_shared<Board>();
    board->board_id_ = board_idx;
    for (int y = 0; y < ny - 1; ++y) for (int x = 0; x < nx - 1; ++x) board->pts_3d_.push_back(Point3f{(float)(x * sq), (float)(y * sq), 0.f});
    boards_3d_[board_idx] = board;
  }
  return true;
}

bool Calibration::boardExtraction() {
  if (keypoints_path_.empty() || keypoints_path_ == "None") { g_last_error = "board detection from images is outside this library: set keypoints_path"; return false; }
  return loadDetectedKeypoints(keypoints_path_);
}

bool Calibration::initializeCalibrationAllCam() {
  if (cam_params_path_.empty() || cam_params_path_ == "None") { g_last_error = "intrinsic initialisation from images (cv::calibrateCamera) is outside this library: set cam_params_path"; return false; }
  return loadInitialIntrinsics(cam_params_path_);
}

void Calibration::initIntrinsic() {
  if (!initializeCalibrationAllCam()) { std::fprintf(stderr, "[McCalibBA] %s\n", g_last_error.c_str()); return; }
  estimatePoseAllBoards();
  if (fix_intrinsic_ == 0) refineIntrinsicAndPoseAllCam();
  computeReproErrAllBoard();
}

bool Calibration::saveIntrinsics(const std::string& path) const {
  std::ofstream f(path);
  if (!f) return false;
  f.precision(17);
  f << "%YAML:1.0\n---\nnb_camera: " << cams_.size() << "\n";
  for (const auto& it : cams_) {
    const auto& in = it.second->intrinsics_;
    f << "camera_" << it.first << ":\n";
    f << "   camera_matrix: !!opencv-matrix\n      rows: 3\n      cols: 3\n      dt: d\n      data: [ " << in[0] << ", 0., " << in[2] << ", 0., " << in[1] << ", " << in[3] << ", 0., 0., 1. ]\n";
    const int nd = it.second->distortion_model_ == 0 ? 5 : 4;
    f << "   distortion_vector: !!opencv-matrix\n      rows: 1\n      cols: " << nd << "\n      dt: d\n      data: [ ";
    for (int k = 0; k < nd; ++k) f << in[4 + k] << (k + 1 < nd ? ", " : " ]\n");
    f << "   distortion_type: " << it.second->distortion_model_ << "\n";
  }
  return (bool)f;
}

}  // namespace McCalib
