// PFM files (second half): the on-disk format of the predicted disparity maps,
// tools/pfm_file_io.py:6-77 (writer call site test_stereo.py:133).  Host-side code of the library: the header
// ("PF" / "Pf", "<width> <height>", "%f" scale, negative = little-endian) followed by the raw fp32 rows.  The
// reference flips the rows in numpy before writing (np.flipud at the call site) and after reading (:44); here
// the flip is an argument, rows are streamed in the requested order without a flipped copy.
#include <errno.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "rsm_common.cuh"

namespace rsm {

void set_io_error(const char* what, const char* path);

static bool host_little_endian() {
  const uint32_t one = 1;
  unsigned char b;
  memcpy(&b, &one, 1);
  return b == 1;
}

struct File {
  FILE* fp = nullptr;
  File(const char* path, const char* mode) : fp(fopen(path, mode)) {}
  ~File() { if (fp) fclose(fp); }
};

}  // namespace rsm

using namespace rsm;

extern "C" int rsm_pfm_write(const char* path, const float* image, int64_t H, int64_t W, int channels, double scale,
                             int flip_rows) {
  if (!path || (!image && H * W > 0)) return RSM_ERR_NULL_POINTER;
  if (H < 0 || W < 0 || (channels != 1 && channels != 3)) return RSM_ERR_INVALID_SHAPE;
  File f(path, "wb");
  if (!f.fp) { set_io_error("cannot open for writing", path); return RSM_ERR_IO; }
  // tools/pfm_file_io.py:66-75: host byte order decides the sign of the scale
  if (host_little_endian()) scale = -scale;
  if (fprintf(f.fp, "%s\n%lld %lld\n%f\n", channels == 3 ? "PF" : "Pf", (long long)W, (long long)H, scale) < 0) {
    set_io_error("write failed", path);
    return RSM_ERR_IO;
  }
  const size_t row = (size_t)W * channels;
  for (int64_t y = 0; y < H && row > 0; ++y) {
    const float* src = image + (size_t)(flip_rows ? H - 1 - y : y) * row;
    if (fwrite(src, sizeof(float), row, f.fp) != row) { set_io_error("write failed", path); return RSM_ERR_IO; }
  }
  if (fflush(f.fp) != 0) { set_io_error("write failed", path); return RSM_ERR_IO; }
  return RSM_OK;
}

extern "C" int rsm_pfm_read_header(const char* path, int64_t* H, int64_t* W, int* channels, double* scale,
                                   int64_t* data_offset) {
  if (!path || !H || !W || !channels || !scale || !data_offset) return RSM_ERR_NULL_POINTER;
  File f(path, "rb");
  if (!f.fp) { set_io_error("cannot open for reading", path); return RSM_ERR_IO; }
  char line[256];
  if (!fgets(line, sizeof(line), f.fp)) { set_io_error("Not a PFM file.", path); return RSM_ERR_IO; }
  line[strcspn(line, "\r\n \t")] = 0;
  if (strcmp(line, "PF") == 0) *channels = 3;
  else if (strcmp(line, "Pf") == 0) *channels = 1;
  else { set_io_error("Not a PFM file.", path); return RSM_ERR_IO; }
  long long w = 0, h = 0;
  if (!fgets(line, sizeof(line), f.fp) || sscanf(line, "%lld %lld", &w, &h) != 2 || w < 0 || h < 0) {
    set_io_error("Malformed PFM header.", path);
    return RSM_ERR_IO;
  }
  if (!fgets(line, sizeof(line), f.fp) || sscanf(line, "%lf", scale) != 1) {
    set_io_error("Malformed PFM header.", path);
    return RSM_ERR_IO;
  }
  *W = w; *H = h;
  *data_offset = (int64_t)ftell(f.fp);
  return RSM_OK;
}

extern "C" int rsm_pfm_read(const char* path, float* image, int64_t H, int64_t W, int channels, int flip_rows) {
  int64_t h = 0, w = 0, off = 0;
  int ch = 0;
  double scale = 0;
  if (int rc = rsm_pfm_read_header(path, &h, &w, &ch, &scale, &off)) return rc;
  if (h != H || w != W || ch != channels) return RSM_ERR_INVALID_SHAPE;
  if (!image && H * W > 0) return RSM_ERR_NULL_POINTER;
  File f(path, "rb");
  if (!f.fp || fseek(f.fp, (long)off, SEEK_SET) != 0) { set_io_error("cannot open for reading", path); return RSM_ERR_IO; }
  const bool swap = (scale < 0) != host_little_endian();      // file byte order differs from the host's
  const size_t row = (size_t)W * channels;
  for (int64_t y = 0; y < H && row > 0; ++y) {
    float* dst = image + (size_t)(flip_rows ? H - 1 - y : y) * row;
    if (fread(dst, sizeof(float), row, f.fp) != row) { set_io_error("file is shorter than its header says", path); return RSM_ERR_IO; }
    if (swap) {
      uint32_t* u = reinterpret_cast<uint32_t*>(dst);
      for (size_t i = 0; i < row; ++i) u[i] = __builtin_bswap32(u[i]);
    }
  }
  return RSM_OK;
}
