// io_main.cpp — CPU-only driver of the host pieces around the refinement path (tests/test_host_io.py):
//   io_main keypoints <boards.txt> <keypoints_in.yml> <cam_params.yml|None> <keypoints_out.yml> <dump.txt>
//       Calibration::loadDetectedKeypoints + loadInitialIntrinsics + saveDetectedKeypoints
//   io_main avgrot <in.txt> <out.txt>          getAverageRotation on lists of rotation vectors
//   io_main grouppose <in.txt> <out.txt>       CameraGroup::computeObjPoseInCameraGroup + CameraGroupObs::computeObjectsPose
//   io_main config <config.yml> <out.txt>      Calibration::initialization (the constructor of the reference's Calibration)
//   io_main yamldump <in.yml> <out.json>       the parsed tree of a cv::FileStorage YAML file as JSON
#include <cstdio>
#include <fstream>
#include <string>

#include "McCalibBA.hpp"
#include "OcvYaml.hpp"

using namespace McCalib;

static int keypoints(int argc, char** argv) {
  if (argc < 7) return 2;
  Calibration calib;
  std::ifstream bf(argv[2]);
  int n_board; bf >> n_board;
  for (int b = 0; b < n_board; ++b) {                       // Board.cpp:64-72: (nx-1)(ny-1) corners at (x sq, y sq, 0)
    int nx, ny; float sq; bf >> nx >> ny >> sq;
    auto board = std::make_shared<Board>();
    board->board_id_ = b;
    for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) board->pts_3d_.push_back(Point3f{x * sq, y * sq, 0.f});
    calib.boards_3d_[b] = board;
  }
  if (!calib.loadDetectedKeypoints(argv[3])) { std::fprintf(stderr, "loadDetectedKeypoints: %s\n", g_last_error.c_str()); return 1; }
  if (std::string(argv[4]) != "None" && !calib.loadInitialIntrinsics(argv[4])) { std::fprintf(stderr, "loadInitialIntrinsics: %s\n", g_last_error.c_str()); return 1; }
  if (!calib.saveDetectedKeypoints(argv[5])) return 1;
  std::ofstream out(argv[6]);
  out.precision(17);
  out << calib.cams_.size() << " " << calib.frames_.size() << " " << calib.board_observations_.size() << "\n";
  for (const auto& c : calib.cams_) {
    out << c.first << " " << c.second->im_cols_ << " " << c.second->im_rows_ << " " << c.second->distortion_model_ << " " << c.second->board_observations_.size();
    for (double x : c.second->intrinsics_) out << " " << x;
    out << "\n";
  }
  for (const auto& it : calib.board_observations_) {
    const BoardObs& bo = *it.second;
    out << bo.camera_id_ << " " << bo.frame_id_ << " " << bo.board_id_ << " " << bo.pts_2d_.size() << " " << (bo.board_3d_.lock() ? 1 : 0);
    for (size_t i = 0; i < bo.pts_2d_.size(); ++i) out << " " << bo.charuco_id_[i] << " " << bo.pts_2d_[i].x << " " << bo.pts_2d_[i].y;
    out << "\n";
  }
  return 0;
}

static int avgrot(char** argv) {
  std::ifstream in(argv[2]); std::ofstream out(argv[3]);
  out.precision(17);
  int n_cases; in >> n_cases;
  for (int c = 0; c < n_cases; ++c) {
    int quat, n; in >> quat >> n;
    std::vector<double> r1(n), r2(n), r3(n);
    for (int i = 0; i < n; ++i) in >> r1[i] >> r2[i] >> r3[i];
    const auto r = getAverageRotation(r1, r2, r3, quat != 0);
    out << r[0] << " " << r[1] << " " << r[2] << "\n";
  }
  return 0;
}

static int grouppose(char** argv) {
  std::ifstream in(argv[2]); std::ofstream out(argv[3]);
  out.precision(17);
  int n_cases; in >> n_cases;
  for (int c = 0; c < n_cases; ++c) {
    int ref_cam, n_cam, n_obs, quat; in >> ref_cam >> n_cam >> n_obs >> quat;
    auto group = std::make_shared<CameraGroup>();
    group->cam_group_idx_ = 0; group->id_ref_cam_ = ref_cam;
    for (int k = 0; k < n_cam; ++k) { Pose6 p; for (double& x : p) in >> x; group->relative_camera_pose_[k] = p; }
    auto frame = std::make_shared<Frame>();
    auto cgo = std::make_shared<CameraGroupObs>();
    cgo->cam_group_idx_ = 0; cgo->cam_group_ = group; cgo->quaternion_averaging_ = quat != 0;
    frame->cam_group_observations_[0] = cgo; group->frames_[0] = frame;
    std::vector<std::shared_ptr<Object3DObs>> keep;
    for (int k = 0; k < n_obs; ++k) {
      auto oo = std::make_shared<Object3DObs>();
      in >> oo->camera_id_ >> oo->object_3d_id_;
      for (double& x : oo->pose_) in >> x;
      keep.push_back(oo); cgo->insertObjectObservation(oo);
    }
    group->computeObjPoseInCameraGroup();
    cgo->computeObjectsPose();
    out << cgo->object_pose_.size();
    for (const auto& it : cgo->object_pose_) { out << " " << it.first; for (double x : it.second) out << " " << x; }
    out << "\n";
    for (const auto& oo : keep) { for (double x : oo->group_pose_) out << x << " "; out << "\n"; }
  }
  return 0;
}

static int config(char** argv) {
  Calibration calib;
  if (!calib.initialization(argv[2])) { std::fprintf(stderr, "initialization: %s\n", g_last_error.c_str()); return 1; }
  std::ofstream out(argv[3]);
  out.precision(9);
  out << calib.cams_.size() << " " << calib.boards_3d_.size() << " " << calib.nb_iterations_ << " " << calib.ransac_thresh_ << " "
      << calib.fix_intrinsic_ << " " << calib.quaternion_averaging_ << "\n";
  out << calib.cam_params_path_ << "\n" << calib.keypoints_path_ << "\n" << calib.save_path_ << "\n";
  for (const auto& c : calib.cams_) out << c.second->distortion_model_ << " ";
  out << "\n";
  for (const auto& b : calib.boards_3d_) {
    out << b.second->pts_3d_.size();
    for (const Point3f& p : b.second->pts_3d_) out << " " << p.x << " " << p.y << " " << p.z;
    out << "\n";
  }
  return 0;
}

// lines "m r0..r8" -> quaternion (x y z w); "q x y z w" -> rotation matrix; "b rx ry rz" (degrees as the reference's test
// passes them to cv::Rodrigues, i.e. used as radians) -> matrix before and after the quaternion round trip
static int quat(char** argv) {
  std::ifstream in(argv[2]); std::ofstream out(argv[3]);
  out.precision(17);
  std::string kind;
  while (in >> kind) {
    if (kind == "m") { double R[9]; for (double& x : R) in >> x; for (double x : convertRotationMatrixToQuaternion(R)) out << x << " "; }
    else if (kind == "q") { std::array<double, 4> q; for (double& x : q) in >> x; for (double x : convertQuaternionToRotationMatrix(q)) out << x << " "; }
    else if (kind == "b") {
      double rv[3]; for (double& x : rv) in >> x;
      double R[9]; rotVecToMat(rv, R);
      const auto R2 = convertQuaternionToRotationMatrix(convertRotationMatrixToQuaternion(R));
      for (double x : R) out << x << " ";
      for (double x : R2) out << x << " ";
    }
    out << "\n";
  }
  return 0;
}

// pose helpers: "r rx ry rz" -> R (9); "v r0..r8" -> rotation vector; "c a(6) b(6)" -> composePose(a, b); "i p(6)" -> invertPose
static int pose(char** argv) {
  std::ifstream in(argv[2]); std::ofstream out(argv[3]);
  out.precision(17);
  std::string kind;
  while (in >> kind) {
    if (kind == "r") { double rv[3], R[9]; for (double& x : rv) in >> x; rotVecToMat(rv, R); for (double x : R) out << x << " "; }
    else if (kind == "v") { double R[9], rv[3]; for (double& x : R) in >> x; matToRotVec(R, rv); for (double x : rv) out << x << " "; }
    else if (kind == "c") { Pose6 a, b; for (double& x : a) in >> x; for (double& x : b) in >> x; for (double x : composePose(a, b)) out << x << " "; }
    else if (kind == "i") { Pose6 a; for (double& x : a) in >> x; for (double x : invertPose(a)) out << x << " "; }
    out << "\n";
  }
  return 0;
}

// canonical JSON of a parsed file: scalars as strings, maps in file order
static void dumpNode(const ocvyaml::Node& n, std::ofstream& out) {
  if (n.kind == ocvyaml::Node::SEQ) {
    out << "[";
    for (size_t i = 0; i < n.seq.size(); ++i) { if (i) out << ","; dumpNode(n.seq[i], out); }
    out << "]";
  } else if (n.kind == ocvyaml::Node::MAP) {
    out << "{";
    for (size_t i = 0; i < n.map.size(); ++i) {
      if (i) out << ",";
      out << "\"" << n.map[i].first << "\":";
      dumpNode(n.map[i].second, out);
    }
    out << "}";
  } else {
    out << "\"";
    for (char c : n.scalar) { if (c == '"' || c == '\\') out << '\\'; out << c; }
    out << "\"";
  }
}
static int yamldump(char** argv) {
  ocvyaml::Node root; ocvyaml::Parser parser; std::string err;
  if (!parser.parseFile(argv[2], root, &err)) { std::fprintf(stderr, "parse: %s\n", err.c_str()); return 1; }
  std::ofstream out(argv[3]);
  dumpNode(root, out);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s keypoints|avgrot|grouppose ...\n", argv[0]); return 2; }
  const std::string mode = argv[1];
  if (mode == "keypoints") return keypoints(argc, argv);
  if (mode == "avgrot") return avgrot(argv);
  if (mode == "grouppose") return grouppose(argv);
  if (mode == "config") return config(argv);
  if (mode == "yamldump") return yamldump(argv);
  if (mode == "quat") return quat(argv);
  if (mode == "pose") return pose(argv);
  return 2;
}
