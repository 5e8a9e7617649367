// calibrate_main.cpp — the head of apps/calibrate/src/calibrate.cpp's runCalibrationWorkflow (:9-18) from files only:
//   McCalib::Calibration Calib(config_path);   -> Calibration::initialization
//   Calib.boardExtraction();                   -> keypoint file (detected_keypoints_data.yml)
//   Calib.initIntrinsic();                     -> camera-parameter file, estimatePoseAllBoards (batched P3P RANSAC on the
//                                                 GPU), refineIntrinsicAndPoseAllCam (batched stage-1 LM on the GPU)
// then writes the refined intrinsics in the camera-parameter layout and, optionally, re-saves the keypoints (the
// outliers removed by the RANSAC are gone from them, as in the reference after estimatePoseAllBoards).
//   usage: calibrate_main <config.yml> <intrinsics_out.yml> [keypoints_out.yml]
#include <cstdio>
#include <string>

#include "McCalibBA.hpp"

using namespace McCalib;

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s config.yml intrinsics_out.yml [keypoints_out.yml]\n", argv[0]); return 2; }
  g_verbose = 0;
  Calibration calib;
  if (!calib.initialization(argv[1])) { std::fprintf(stderr, "config: %s\n", g_last_error.c_str()); return 1; }
  if (!calib.boardExtraction()) { std::fprintf(stderr, "boardExtraction: %s\n", g_last_error.c_str()); return 1; }
  std::printf("cameras %zu boards %zu frames %zu board observations %zu\n", calib.cams_.size(), calib.boards_3d_.size(),
              calib.frames_.size(), calib.board_observations_.size());
  g_last_error.clear();
  calib.initIntrinsic();
  if (!g_last_error.empty()) { std::fprintf(stderr, "initIntrinsic: %s\n", g_last_error.c_str()); return 1; }
  size_t valid = 0, corners = 0;
  for (const auto& it : calib.board_observations_) { valid += it.second->valid_ ? 1 : 0; corners += it.second->pts_2d_.size(); }
  std::printf("valid board observations %zu corners kept %zu mean board reprojection error %.6f px\n", valid, corners,
              calib.computeReproErrAllBoard());
  if (!calib.saveIntrinsics(argv[2])) { std::fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
  if (argc > 3 && !calib.saveDetectedKeypoints(argv[3])) { std::fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
  return 0;
}
