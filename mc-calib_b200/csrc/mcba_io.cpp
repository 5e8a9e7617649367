// mcba_io.cpp — binary dump / load of a packed problem and its parameter blocks (checkpoint / resume row).
//
// The reference's de-facto checkpoint is the pair detected_keypoints_data.yml + camera-parameter file
// (McCalib/src/McCalib.cpp:493-548 writes, :172-226 and :662-688 reload): text, per frame and per board, and it stops
// before the problem the refine* stages build. This is the same idea one step later: the flat arrays mcba_create takes
// (include/mcba.h, mcba_problem) and the four parameter arrays (mcba_params), as they are, so that a bundle adjustment
// can be replayed, resumed from the parameters of an earlier solve, or shipped as a parity fixture without the
// generator that made it. Host code only; nothing here touches the device.
//
// Layout (little endian, no padding between sections, every section 8-byte aligned by construction of the header):
//   char    magic[8] = "MCBAPRB1"
//   int32   stage, n_cam, n_pose, n_board, n_pts, n_problems, has (bit 0 cam_is_ref, 1 board_is_ref, 2 cam_problem,
//           3 params.cam_pose, 4 params.pose, 5 params.board_pose, 6 params.intrinsics), reserved
//   int64   n_obs, n_bobs
//   u[n_obs] v[n_obs] (f32) | pt_idx[n_obs] (i32) | bobs_ptr[n_bobs+1] (i64) | bobs_cam, bobs_pose, bobs_board [n_bobs] (i32)
//   | pts_xyz[3 n_pts] (f32) | cam_model[n_cam] (i32) | cam_is_ref[n_cam] | board_is_ref[n_board] (u8, if present)
//   | cam_problem[n_cam] (i32, if present) | cam_pose[6 n_cam] pose[6 n_pose] board_pose[6 n_board] intrinsics[9 n_cam]
//   (f64, if present); each section padded with zeros to a multiple of 8 bytes.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/mcba.h"

namespace {

thread_local std::string g_io_err;

struct Storage {            // owns everything a loaded problem points to
  std::vector<std::vector<char>> blocks;
  void* add(size_t bytes) { blocks.emplace_back(bytes ? bytes : 1); return blocks.back().data(); }
};

struct Header {
  char magic[8];
  int32_t stage, n_cam, n_pose, n_board, n_pts, n_problems, has, reserved;
  int64_t n_obs, n_bobs;
};
static_assert(sizeof(Header) == 56, "header layout");

const char MAGIC[8] = {'M', 'C', 'B', 'A', 'P', 'R', 'B', '1'};

bool put(FILE* f, const void* p, size_t bytes) {
  static const char zero[8] = {0};
  if (bytes && fwrite(p, 1, bytes, f) != bytes) return false;
  const size_t pad = (8 - bytes % 8) % 8;
  return pad == 0 || fwrite(zero, 1, pad, f) == pad;
}
bool get(FILE* f, void* p, size_t bytes) {
  char skip[8];
  if (bytes && fread(p, 1, bytes, f) != bytes) return false;
  const size_t pad = (8 - bytes % 8) % 8;
  return pad == 0 || fread(skip, 1, pad, f) == pad;
}

}  // namespace

extern "C" {

const char* mcba_io_last_error(void) { return g_io_err.c_str(); }

int mcba_problem_save(const char* path, const mcba_problem* P, const mcba_params* X) {
  if (!path || !P) { g_io_err = "null argument"; return MCBA_ERR_INVALID; }
  if (P->n_obs < 0 || P->n_bobs < 0 || P->n_cam < 0 || P->n_pose < 0 || P->n_board < 0 || P->n_pts < 0) { g_io_err = "negative count"; return MCBA_ERR_INVALID; }
  if ((P->n_obs && (!P->u || !P->v || !P->pt_idx)) || !P->bobs_ptr || (P->n_bobs && (!P->bobs_cam || !P->bobs_pose)) ||
      (P->n_pts && !P->pts_xyz) || (P->n_cam && !P->cam_model)) { g_io_err = "missing array"; return MCBA_ERR_INVALID; }
  FILE* f = fopen(path, "wb");
  if (!f) { g_io_err = std::string("cannot open ") + path; return MCBA_ERR_INVALID; }
  Header h{};
  memcpy(h.magic, MAGIC, 8);
  h.stage = P->stage; h.n_cam = P->n_cam; h.n_pose = P->n_pose; h.n_board = P->n_board; h.n_pts = P->n_pts;
  h.n_problems = P->n_problems; h.n_obs = P->n_obs; h.n_bobs = P->n_bobs;
  h.has = (P->cam_is_ref ? 1 : 0) | (P->board_is_ref ? 2 : 0) | (P->cam_problem ? 4 : 0) |
          ((X && X->cam_pose) ? 8 : 0) | ((X && X->pose) ? 16 : 0) | ((X && X->board_pose) ? 32 : 0) | ((X && X->intrinsics) ? 64 : 0);
  const size_t no = (size_t)P->n_obs, nb = (size_t)P->n_bobs, nc = (size_t)P->n_cam;
  std::vector<int32_t> zeros_b;
  const int32_t* board = P->bobs_board;
  if (!board) { zeros_b.assign(nb, 0); board = zeros_b.data(); }   // ignored in stages 1 and 3: stored as zeros
  bool ok = put(f, &h, sizeof h) && put(f, P->u, 4 * no) && put(f, P->v, 4 * no) && put(f, P->pt_idx, 4 * no) &&
            put(f, P->bobs_ptr, 8 * (nb + 1)) && put(f, P->bobs_cam, 4 * nb) && put(f, P->bobs_pose, 4 * nb) &&
            put(f, board, 4 * nb) && put(f, P->pts_xyz, 12 * (size_t)P->n_pts) && put(f, P->cam_model, 4 * nc);
  if (ok && (h.has & 1)) ok = put(f, P->cam_is_ref, nc);
  if (ok && (h.has & 2)) ok = put(f, P->board_is_ref, (size_t)P->n_board);
  if (ok && (h.has & 4)) ok = put(f, P->cam_problem, 4 * nc);
  if (ok && (h.has & 8)) ok = put(f, X->cam_pose, 48 * nc);
  if (ok && (h.has & 16)) ok = put(f, X->pose, 48 * (size_t)P->n_pose);
  if (ok && (h.has & 32)) ok = put(f, X->board_pose, 48 * (size_t)P->n_board);
  if (ok && (h.has & 64)) ok = put(f, X->intrinsics, 72 * nc);
  if (fclose(f) != 0) ok = false;
  if (!ok) { g_io_err = std::string("write error on ") + path; return MCBA_ERR_INVALID; }
  return MCBA_OK;
}

int mcba_problem_load(const char* path, mcba_problem* P, mcba_params* X, void** storage) {
  if (!path || !P || !storage) { g_io_err = "null argument"; return MCBA_ERR_INVALID; }
  *storage = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { g_io_err = std::string("cannot open ") + path; return MCBA_ERR_INVALID; }
  Header h;
  if (fread(&h, 1, sizeof h, f) != sizeof h || memcmp(h.magic, MAGIC, 8) != 0) { fclose(f); g_io_err = "not an MCBAPRB1 file"; return MCBA_ERR_INVALID; }
  if (h.n_obs < 0 || h.n_bobs < 0 || h.n_cam < 0 || h.n_pose < 0 || h.n_board < 0 || h.n_pts < 0 || h.stage < 1 || h.stage > 5) {
    fclose(f); g_io_err = "corrupt header"; return MCBA_ERR_INVALID;
  }
  // the counts must be consistent with the file size before anything is allocated from them
  {
    const size_t no = (size_t)h.n_obs, nb = (size_t)h.n_bobs, nc = (size_t)h.n_cam;
    auto r8 = [](size_t b) { return (b + 7) / 8 * 8; };
    size_t need = sizeof h + 3 * r8(4 * no) + r8(8 * (nb + 1)) + 3 * r8(4 * nb) + r8(12 * (size_t)h.n_pts) + r8(4 * nc);
    if (h.has & 1) need += r8(nc);
    if (h.has & 2) need += r8((size_t)h.n_board);
    if (h.has & 4) need += r8(4 * nc);
    if (h.has & 8) need += 48 * nc;
    if (h.has & 16) need += 48 * (size_t)h.n_pose;
    if (h.has & 32) need += 48 * (size_t)h.n_board;
    if (h.has & 64) need += 72 * nc;
    fseek(f, 0, SEEK_END);
    const long have = ftell(f);
    fseek(f, (long)sizeof h, SEEK_SET);
    if (have < 0 || (size_t)have != need) { fclose(f); g_io_err = "file size does not match its header"; return MCBA_ERR_INVALID; }
  }
  Storage* S = new Storage;
  const size_t no = (size_t)h.n_obs, nb = (size_t)h.n_bobs, nc = (size_t)h.n_cam;
  bool ok = true;
  auto rd = [&](size_t bytes) -> void* { void* p = S->add(bytes); if (ok) ok = get(f, p, bytes); return p; };
  memset(P, 0, sizeof *P);
  P->stage = h.stage; P->n_cam = h.n_cam; P->n_pose = h.n_pose; P->n_board = h.n_board; P->n_pts = h.n_pts;
  P->n_problems = h.n_problems; P->n_obs = h.n_obs; P->n_bobs = h.n_bobs;
  P->u = (const float*)rd(4 * no); P->v = (const float*)rd(4 * no); P->pt_idx = (const int32_t*)rd(4 * no);
  P->bobs_ptr = (const int64_t*)rd(8 * (nb + 1));
  P->bobs_cam = (const int32_t*)rd(4 * nb); P->bobs_pose = (const int32_t*)rd(4 * nb); P->bobs_board = (const int32_t*)rd(4 * nb);
  P->pts_xyz = (const float*)rd(12 * (size_t)h.n_pts); P->cam_model = (const int32_t*)rd(4 * nc);
  if (h.has & 1) P->cam_is_ref = (const uint8_t*)rd(nc);
  if (h.has & 2) P->board_is_ref = (const uint8_t*)rd((size_t)h.n_board);
  if (h.has & 4) P->cam_problem = (const int32_t*)rd(4 * nc);
  mcba_params x{};
  if (h.has & 8) x.cam_pose = (double*)rd(48 * nc);
  if (h.has & 16) x.pose = (double*)rd(48 * (size_t)h.n_pose);
  if (h.has & 32) x.board_pose = (double*)rd(48 * (size_t)h.n_board);
  if (h.has & 64) x.intrinsics = (double*)rd(72 * nc);
  fclose(f);
  if (!ok) { delete S; g_io_err = "truncated file"; return MCBA_ERR_INVALID; }
  if (X) *X = x;
  *storage = S;
  return MCBA_OK;
}

void mcba_problem_free(void* storage) { delete (Storage*)storage; }

}  // extern "C"
