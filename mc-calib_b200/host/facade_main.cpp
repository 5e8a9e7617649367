// facade_main.cpp — drives the McCalib-shaped facade end to end on a scene file written by tests/test_gpu_facade.py:
// builds the Camera / Board / Object3D / Frame / CameraGroup graph the way MC-Calib's pipeline leaves it just before
// the final refinement, runs the requested refine* entry points of Calibration, and prints the refined blocks.
//   usage: facade_main <scene.txt> <flow> <out.txt>     flow: "final" (apps/calibrate/src/calibrate.cpp:60-64),
//                                                             "intrinsics" (refineIntrinsicAndPoseAllCam),
//                                                             "objects" (refineAllObject3D), "group" (refineAllCameraGroup),
//                                                             "init" (no pose is taken from the file: estimatePoseAllBoards,
//                                                             estimatePoseAllObjects, computeAllObjPoseInCameraGroup on the
//                                                             GPU / host), "init_final" ("init" followed by "final")
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "McCalibBA.hpp"

using namespace McCalib;

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s scene.txt flow out.txt\n", argv[0]); return 2; }
  std::ifstream in(argv[1]);
  if (!in) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  const std::string flow = argv[2];
  Calibration calib;
  g_verbose = 0;
  if (const char* d = std::getenv("MCBA_DUMP_PROBLEM_DIR")) g_problem_dump_dir = d;   // checkpoint every packed problem (mcba_problem_save)
  int n_cam, n_board, n_frames, n_bobs;
  in >> n_cam >> n_board >> n_frames >> n_bobs >> calib.nb_iterations_ >> calib.fix_intrinsic_;
  const bool do_init = flow == "init" || flow == "init_final";
  if (argc > 4) calib.ransac_thresh_ = (float)std::atof(argv[4]);
  if (argc > 5) calib.ransac_seed_ = std::strtoull(argv[5], nullptr, 10);
  auto group = std::make_shared<CameraGroup>();
  group->cam_group_idx_ = 0; group->id_ref_cam_ = 0;
  calib.cam_group_[0] = group;
  for (int c = 0; c < n_cam; ++c) {
    auto cam = std::make_shared<Camera>();
    cam->cam_idx_ = c;
    in >> cam->distortion_model_;
    for (double& x : cam->intrinsics_) in >> x;
    Pose6 p; for (double& x : p) in >> x;
    calib.cams_[c] = cam; group->cameras_[c] = cam; group->relative_camera_pose_[c] = p;
  }
  auto object = std::make_shared<Object3D>();
  object->obj_id_ = 0; object->ref_board_id_ = 0;
  calib.object_3d_[0] = object;
  for (int b = 0; b < n_board; ++b) {
    auto board = std::make_shared<Board>();
    board->board_id_ = b;
    int npts; in >> npts;
    board->pts_3d_.resize(npts);
    for (auto& p : board->pts_3d_) in >> p.x >> p.y >> p.z;
    Pose6 p; for (double& x : p) in >> x;
    calib.boards_3d_[b] = board; object->boards_[b] = board; object->relative_board_pose_[b] = p;
    for (int i = 0; i < npts; ++i) {
      object->pts_board_2_obj_[std::make_pair(b, i)] = (int)object->pts_obj_2_board_.size();
      object->pts_obj_2_board_.push_back(std::make_pair(b, i));
    }
  }
  object->pts_3d_.resize(object->pts_obj_2_board_.size());
  object->updateObjectPts();
  for (int f = 0; f < n_frames; ++f) {
    auto frame = std::make_shared<Frame>();
    frame->frame_idx_ = f;
    auto cgo = std::make_shared<CameraGroupObs>();
    cgo->cam_group_idx_ = 0; cgo->cam_group_ = group;
    Pose6 p; for (double& x : p) in >> x;
    if (do_init) p = Pose6{};                 // to be computed from the observations
    cgo->object_pose_[0] = p;
    calib.frames_[f] = frame; calib.cams_group_obs_[f] = cgo;
    frame->cam_group_observations_[0] = cgo; group->frames_[f] = frame;
  }
  // board observations: frame cam board npts (charuco_id u v)*, sorted by (frame, cam, board)
  int n_bo = 0, n_oo = 0, last_frame = -1, last_cam = -1;
  for (int k = 0; k < n_bobs; ++k) {
    auto bo = std::make_shared<BoardObs>();
    int npts;
    in >> bo->frame_id_ >> bo->camera_id_ >> bo->board_id_ >> npts;
    bo->pts_2d_.resize(npts); bo->charuco_id_.resize(npts);
    for (int i = 0; i < npts; ++i) in >> bo->charuco_id_[i] >> bo->pts_2d_[i].x >> bo->pts_2d_[i].y;
    bo->cam_ = calib.cams_[bo->camera_id_]; bo->board_3d_ = calib.boards_3d_[bo->board_id_];
    calib.board_observations_[n_bo] = bo;
    calib.cams_[bo->camera_id_]->board_observations_[n_bo] = bo;
    ++n_bo;
  }
  if (do_init) calib.estimatePoseAllBoards();   // McCalib.cpp:703-706: BoardObs::pose_, valid_, inlier corners only
  // object observations, from the (valid) board observations of one (frame, camera) (Object3DObs.cpp:31-52)
  std::shared_ptr<Object3DObs> oo;
  for (int k = 0; k < n_bo; ++k) {
    auto bo = calib.board_observations_[k];
    if (!bo->valid_) continue;
    auto cgo = calib.cams_group_obs_[bo->frame_id_];
    if (!do_init) {
      // BoardObs::pose_ = camera o object o board, Object3DObs::pose_ = camera o object (what estimatePose* would leave)
      const Pose6 cam_obj = composePose(group->relative_camera_pose_[bo->camera_id_], cgo->object_pose_[0]);
      bo->pose_ = composePose(cam_obj, object->relative_board_pose_[bo->board_id_]);
    }
    if (bo->frame_id_ != last_frame || bo->camera_id_ != last_cam) {
      oo = std::make_shared<Object3DObs>();
      oo->frame_id_ = bo->frame_id_; oo->camera_id_ = bo->camera_id_; oo->object_3d_id_ = 0;
      oo->cam_ = calib.cams_[bo->camera_id_]; oo->object_3d_ = object;
      if (!do_init) oo->pose_ = composePose(group->relative_camera_pose_[bo->camera_id_], cgo->object_pose_[0]);
      calib.object_observations_[n_oo] = oo; object->object_observations_[n_oo] = oo;
      cgo->insertObjectObservation(oo);
      ++n_oo; last_frame = bo->frame_id_; last_cam = bo->camera_id_;
    }
    oo->board_observations_[k] = bo;
    for (size_t i = 0; i < bo->charuco_id_.size(); ++i) {   // corners appended board after board
      oo->pts_2d_.push_back(bo->pts_2d_[i]);
      oo->pts_id_.push_back(object->pts_board_2_obj_[std::make_pair(bo->board_id_, bo->charuco_id_[i])]);
    }
  }
  if (!in) { std::fprintf(stderr, "scene file truncated\n"); return 2; }

  if (do_init) {
    calib.estimatePoseAllObjects();             // McCalib.cpp:990-993
    calib.computeAllObjPoseInCameraGroup();     // McCalib.cpp:1635-1641
    if (!g_last_error.empty()) { std::fprintf(stderr, "libmcba failed: %s\n", g_last_error.c_str()); return 1; }
    std::ofstream io(std::string(argv[3]) + ".init.txt");
    io.precision(17);
    for (int k = 0; k < n_bo; ++k) io << (calib.board_observations_[k]->valid_ ? 1 : 0) << " " << calib.board_observations_[k]->charuco_id_.size() << "\n";
    for (int k = 0; k < n_oo; ++k) { io << calib.object_observations_[k]->pts_id_.size(); for (double x : calib.object_observations_[k]->group_pose_) io << " " << x; io << "\n"; }
  }
  if (flow == "final" || flow == "init_final") {   // apps/calibrate/src/calibrate.cpp:60-64
    calib.refineAllCameraGroupAndObjects();
    if (calib.fix_intrinsic_ == 0) calib.refineAllCameraGroupAndObjectsAndIntrinsic();
  } else if (flow == "intrinsics") calib.refineIntrinsicAndPoseAllCam();
  else if (flow == "objects") calib.refineAllObject3D();
  else if (flow == "group") calib.refineAllCameraGroup();
  else if (flow == "init") {}
  else { std::fprintf(stderr, "unknown flow %s\n", flow.c_str()); return 2; }
  if (!g_last_error.empty()) { std::fprintf(stderr, "libmcba failed: %s\n", g_last_error.c_str()); return 1; }
  const double avg_err = calib.computeAvgReprojectionError();

  std::ofstream out(argv[3]);
  out.precision(17);
  out << avg_err << "\n";
  for (int c = 0; c < n_cam; ++c) { for (double x : calib.cams_[c]->intrinsics_) out << x << " "; for (double x : group->relative_camera_pose_[c]) out << x << " "; out << "\n"; }
  for (int b = 0; b < n_board; ++b) { for (double x : object->relative_board_pose_[b]) out << x << " "; out << "\n"; }
  for (int f = 0; f < n_frames; ++f) { for (double x : calib.cams_group_obs_[f]->object_pose_[0]) out << x << " "; out << "\n"; }
  for (int k = 0; k < n_bo; ++k) { for (double x : calib.board_observations_[k]->pose_) out << x << " "; out << "\n"; }
  for (int k = 0; k < n_oo; ++k) { for (double x : calib.object_observations_[k]->pose_) out << x << " "; out << "\n"; }
  calib.saveCamerasParams(std::string(argv[3]) + ".cameras.yml");
  calib.save3DObj(std::string(argv[3]) + ".objects.yml");
  calib.save3DObjPose(std::string(argv[3]) + ".objects_pose.yml");
  calib.saveReprojectionErrorToFile(std::string(argv[3]) + ".reprojection_error.yml");
  {   // Calibration::reproErrorAllCamGroup (apps/calibrate/src/calibrate.cpp:66): mean error per object observation
    std::ofstream ro(std::string(argv[3]) + ".repro_obs.txt");
    ro.precision(9);
    for (const auto& r : calib.reproErrorAllCamGroup())
      ro << r.group << " " << r.frame << " " << r.camera << " " << r.object << " " << r.mean_error << " " << r.nb_pts << "\n";
  }
  return 0;
}
