// McCalibBA.hpp — host-side mirror of the MC-Calib classes that own the bundle-adjustment state, with the
// reference's refinement entry points re-implemented on top of libmcba (include/mcba.h).
//
// Mirrors (names, members and call order; /root/reference/McCalib/include/*.hpp):
//   Camera.hpp:43-46 (intrinsics_), Board.hpp:36 (pts_3d_), BoardObs.hpp:33-37 (pose_, pts_2d_, charuco_id_),
//   Object3D.hpp:43-48 (pts_obj_2_board_, relative_board_pose_), Object3DObs.hpp:40-47 (pose_, pts_2d_, pts_id_),
//   CameraGroupObs.hpp:29-30 (object_pose_), CameraGroup.hpp:40-41 (relative_camera_pose_), Frame.hpp,
//   McCalib.hpp (Calibration: refineIntrinsicAndPoseAllCam, refineAllObject3D, refineAllCameraGroup,
//   refineAllCameraGroupAndObjects, refineAllCameraGroupAndObjectsAndIntrinsic, computeAvgReprojectionError).
// Only what the refinement path reads or writes is mirrored: detection, initialisation, graphs and file I/O of
// the reference are out of scope. OpenCV types are replaced by plain structs.
#pragma once
#include <array>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace McCalib {

struct Point2f { float x, y; };
struct Point3f { float x, y, z; };
typedef std::array<double, 6> Pose6;   // angle-axis (3) || translation (3)

class Board; class BoardObs; class Camera; class Object3D; class Object3DObs; class CameraGroup; class CameraGroupObs; class Frame;

class Board final {
public:
  int board_id_ = 0;
  std::vector<Point3f> pts_3d_;   // (x*sq, y*sq, 0), Board.cpp:66-72
};

class Camera final {
public:
  int cam_idx_ = 0;
  int distortion_model_ = 0;              // 0 Brown, 1 Kannala
  int im_cols_ = 0, im_rows_ = 0;         // Camera.hpp (img_width / img_height of the keypoint file)
  std::array<double, 9> intrinsics_{};    // fx fy u0 v0 d4..d8
  std::map<int, std::weak_ptr<BoardObs>> board_observations_;
  void refineIntrinsicCalibration(const int nb_iterations);   // Camera.cpp:268-305
};

class BoardObs final {
public:
  int frame_id_ = 0, camera_id_ = 0, board_id_ = 0;
  std::vector<Point2f> pts_2d_;
  std::vector<int> charuco_id_;
  Pose6 pose_{};          // board pose w.r.t. the camera
  bool valid_ = true;
  std::weak_ptr<Camera> cam_;
  std::weak_ptr<Board> board_3d_;
  void estimatePose(const float ransac_thresh, const int ransac_iterations);   // BoardObs.cpp:121-163 (GPU, one observation)
};

class Object3D final {
public:
  int obj_id_ = 0, ref_board_id_ = 0;
  std::map<int, std::weak_ptr<Board>> boards_;
  std::map<int, Pose6> relative_board_pose_;
  std::vector<Point3f> pts_3d_;                                  // object-frame points (float, regenerated)
  std::map<std::pair<int, int>, int> pts_board_2_obj_;
  std::vector<std::pair<int, int>> pts_obj_2_board_;
  std::map<int, std::weak_ptr<Object3DObs>> object_observations_;
  void refineObject(const int nb_iterations);   // Object3D.cpp:160-243
  void updateObjectPts();                       // Object3D.cpp:249-268
};

class Object3DObs final {
public:
  int frame_id_ = 0, camera_id_ = 0, object_3d_id_ = 0;
  std::vector<Point2f> pts_2d_;
  std::vector<int> pts_id_;                 // indices into Object3D::pts_3d_
  Pose6 pose_{};                            // object pose w.r.t. the camera
  Pose6 group_pose_{};                      // object pose w.r.t. the camera group (setPoseInGroupVec)
  std::map<int, std::weak_ptr<BoardObs>> board_observations_;
  std::weak_ptr<Camera> cam_;
  std::weak_ptr<Object3D> object_3d_;
  void estimatePose(const float ransac_thresh, const int ransac_iterations);   // Object3DObs.cpp:222-262
};

class CameraGroupObs final {
public:
  int cam_group_idx_ = 0;
  bool quaternion_averaging_ = true;        // CameraGroupObs.hpp
  std::weak_ptr<CameraGroup> cam_group_;
  std::vector<int> object_idx_;             // object id of object_observations_[k] (insertObjectObservation order)
  std::map<int, std::weak_ptr<Object3DObs>> object_observations_;
  std::map<int, Pose6> object_pose_;        // object pose w.r.t. the reference camera of the group
  void insertObjectObservation(const std::shared_ptr<Object3DObs>& obs);   // CameraGroupObs.cpp:26-33
  void computeObjectsPose();                // CameraGroupObs.cpp:42-108
  void updateObjObsPose();                  // CameraGroupObs.cpp:208-217
};

class Frame final {
public:
  int frame_idx_ = 0;
  std::map<int, std::string> frame_path_;   // per camera (Frame.hpp)
  std::map<int, std::weak_ptr<BoardObs>> board_observations_;
  std::map<int, std::weak_ptr<CameraGroupObs>> cam_group_observations_;
};

class CameraGroup final {
public:
  int cam_group_idx_ = 0, id_ref_cam_ = 0;
  std::map<int, std::weak_ptr<Camera>> cameras_;
  std::map<int, std::weak_ptr<Frame>> frames_;
  std::map<int, Pose6> relative_camera_pose_;
  void computeObjPoseInCameraGroup();                                      // CameraGroup.cpp:153-183
  void refineCameraGroup(const int nb_iterations);                         // CameraGroup.cpp:191-279
  void refineCameraGroupAndObjects(const int nb_iterations);               // CameraGroup.cpp:366-474
  void refineCameraGroupAndObjectsAndIntrinsics(const int nb_iterations);  // CameraGroup.cpp:483-583
  // stage 4 immediately followed by stage 5 on one device-resident copy of the observations
  // (apps/calibrate/src/calibrate.cpp:60-64 calls the two entry points back to back)
  void refineCameraGroupAndObjectsThenIntrinsics(const int nb_iterations, bool with_intrinsics);
private:
  void refineStage(int stage, int nb_iterations, bool then_intrinsics);
};

class Calibration final {
public:
  int nb_iterations_ = 1000;       // YAML number_iterations (McCalib.cpp:59)
  int fix_intrinsic_ = 0;          // YAML fix_intrinsic     (McCalib.cpp:77)
  int distortion_model_ = 0;       // YAML distortion_model  (McCalib.cpp:60)
  std::vector<int> distortion_per_camera_;
  int cuda_device_ = 0;
  float ransac_thresh_ = 10.f;     // YAML ransac_threshold  (McCalib.cpp:58)
  unsigned long long ransac_seed_ = 0;   // the reference draws from std::random_device; here the draws are seeded
  std::map<int, std::shared_ptr<Camera>> cams_;
  std::map<int, std::shared_ptr<Board>> boards_3d_;
  std::map<int, std::shared_ptr<BoardObs>> board_observations_;
  std::map<int, std::shared_ptr<Object3D>> object_3d_;
  std::map<int, std::shared_ptr<Object3DObs>> object_observations_;
  std::map<int, std::shared_ptr<CameraGroup>> cam_group_;
  std::map<int, std::shared_ptr<CameraGroupObs>> cams_group_obs_;
  std::map<int, std::shared_ptr<Frame>> frames_;

  bool loadConfig(const std::string& yaml_path);       // the keys of McCalib.cpp:39-79 that the refinement path uses
  // Calibration::Calibration(config_path) (McCalib.cpp:30-135): the keys above plus number_camera, number_board,
  // boards_index, number_x/y_square[_per_board], square_size[_per_board], quaternion_averaging, cam_params_path,
  // keypoints_path, save_path; creates cams_ (Camera(i, distortion)) and boards_3d_ (Board.cpp:22-72 corner grid)
  bool initialization(const std::string& config_path);
  std::string cam_params_path_, keypoints_path_, save_path_;
  int quaternion_averaging_ = 1;
  bool boardExtraction();                              // McCalib.cpp:219-226, keypoint-file branch (detection is out of scope)
  bool initializeCalibrationAllCam();                  // McCalib.cpp:662-695, parameter-file branch (cv::calibrateCamera is out of scope)
  void initIntrinsic();                                // McCalib.cpp:2085-2092
  double computeReproErrAllBoard();                    // McCalib.cpp:722-726 (returns the mean of the per-board means it logs)
  bool saveIntrinsics(const std::string& path) const;  // camera_matrix / distortion_vector / distortion_type per camera: the
                                                       // layout initializeCalibrationAllCam reads back
  // on-disk formats either side of the path (cv::FileStorage YAML)
  void insertNewBoard(const int cam_idx, const int frame_idx, const int board_idx, const std::vector<Point2f>& pts_2d,
                      const std::vector<int>& charuco_idx, const std::string& frame_path);   // McCalib.cpp:594-642
  bool loadDetectedKeypoints(const std::string& path);          // McCalib.cpp:172-213 (detected_keypoints_data.yml)
  bool saveDetectedKeypoints(const std::string& path) const;    // McCalib.cpp:493-548
  bool loadInitialIntrinsics(const std::string& path);          // McCalib.cpp:662-688 (file branch of initializeCalibrationAllCam)
  // pose initialisation feeding the refinement: every observation in one GPU launch (include/mcba_pnp.h)
  void estimatePoseAllBoards();                        // McCalib.cpp:703-706
  void estimatePoseAllObjects();                       // McCalib.cpp:990-993
  void computeAllObjPoseInCameraGroup();               // McCalib.cpp:1198-1205: per group computeObjPoseInCameraGroup, per
                                                       // group observation computeObjectsPose
  void refineIntrinsicAndPoseAllCam();                 // McCalib.cpp:713-716
  void refineAllObject3D();                            // McCalib.cpp:1014-1017
  void refineAllCameraGroup();                         // McCalib.cpp:1211-1220
  void refineAllCameraGroupAndObjects();               // McCalib.cpp:1876-1887
  void refineAllCameraGroupAndObjectsAndIntrinsic();   // McCalib.cpp:2348-2359
  double computeAvgReprojectionError();                // McCalib.cpp:2168-2245 (mean of per-observation means)
  struct ObsMeanError { int group, frame, camera, object; float mean_error; int nb_pts; };
  std::vector<ObsMeanError> reproErrorAllCamGroup();   // McCalib.cpp:1866-1869 -> CameraGroup::reproErrorCameraGroup (CameraGroup.cpp:285-358)
  bool saveCamerasParams(const std::string& path) const;   // calibrated_cameras_data.yml layout of McCalib.cpp:374-404
  bool save3DObj(const std::string& path) const;           // calibrated_objects_data.yml, McCalib.cpp:406-446
  bool save3DObjPose(const std::string& path) const;       // calibrated_objects_pose_data.yml, McCalib.cpp:452-483
  bool saveReprojectionErrorToFile(const std::string& path);   // reprojection_error_data.yml, McCalib.cpp:2251-2341
};

// pose helpers (Rodrigues both ways, composition), double precision
void rotVecToMat(const double* rv, double* R /*9, row-major*/);
void matToRotVec(const double* R, double* rv);
Pose6 composePose(const Pose6& outer, const Pose6& inner);   // x -> outer(inner(x))
Pose6 invertPose(const Pose6& p);
// geometrytools.cpp:832-893: Markley quaternion averaging (largest eigenvector of the mean outer product, antipodal
// quaternions flipped to w >= 0) or the per-component median of the rotation vectors
std::array<double, 4> convertRotationMatrixToQuaternion(const double* R /*9, row-major*/);   // geometrytools.cpp:771-806, (x, y, z, w)
std::array<double, 9> convertQuaternionToRotationMatrix(const std::array<double, 4>& q);     // geometrytools.cpp:808-830
std::array<double, 3> getAverageRotation(std::vector<double>& r1, std::vector<double>& r2, std::vector<double>& r3,
                                         const bool use_quaternion_averaging = true);

extern int g_verbose;              // 1: per-iteration table on stdout like minimizer_progress_to_stdout (default)
extern std::string g_problem_dump_dir;   // non-empty: solveStages writes every packed problem + start parameters there (MCBAPRB1, include/mcba.h)
extern int g_cuda_device;          // device used by the refine* methods (set from Calibration::cuda_device_)
extern std::string g_last_error;   // text of the last libmcba failure (the reference ignores solver failures)

}  // namespace McCalib
