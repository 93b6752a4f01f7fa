// sensit_io.cu -- the reference's on-disk sensitivity formats (sensit.readFromFiles = 1 keeps working).
//
//   sensit_<grav|magn>_<nbproc>_<rank>   big-endian stream (the reference is built with -fconvert=big-endian,
//                                        Makefile:51): header 5 x int32 (ndata_loc, ndata, nelements_total, myrank,
//                                        nbproc) (sensitivity_gravmag.F90:183), then per (station, d, k):
//                                        4 x int32 (idata, nel, k, d) + nel x int32 columns + nel x real(4) values
//                                        (:306-309). Columns are 1-based cells, values carry NO problem / data weight.
//   sensit_<..>_meta.txt                 5 list-directed text lines (:366-374)
//   sensit_<..>_nnz                      int32 N + N x int32 (:386-391)
//   sensit_<..>_weight                   int32 N + N x real(8) (:455-460)
//
// Writer: the rows come from the device-resident row shard (tfx_sensit_assemble_rows). Reader: every rank scans
// the files of all writer ranks and keeps the columns of its own slab (read_sensitivity_kernel, :648-883, with the
// rank-0 read + per-row MPI_Scatterv replaced by independent reads), applies the index shift (:834) and the real(4)
// weights (:837-843) and builds the device matrix. Host integer / byte work; the GPU only receives the result.
#include "../../include/tfx.h"

#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

namespace tfx {
namespace {

const char *suffix(int problem_type) { return problem_type == 2 ? "magn" : "grav"; }   // :57

std::string join(const char *dir, const std::string &name) {
  std::string d = dir ? dir : ".";
  if (!d.empty() && d.back() != '/') d += '/';
  return d + name;
}
std::string rank_file(const char *dir, int problem_type, int nbproc, int rank) {
  return join(dir, std::string("sensit_") + suffix(problem_type) + "_" + std::to_string(nbproc) + "_" + std::to_string(rank));
}

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

// Buffered big-endian writer / reader of 4- and 8-byte items.
struct BeFile {
  FILE *f = nullptr;
  ~BeFile() { if (f) fclose(f); }
  bool open(const std::string &path, const char *mode) { f = fopen(path.c_str(), mode); return f != nullptr; }
  bool put32(const void *src, size_t n) {
    const uint32_t *s = (const uint32_t *)src;
    uint32_t buf[4096];
    while (n) {
      const size_t c = std::min<size_t>(n, 4096);
      for (size_t i = 0; i < c; ++i) buf[i] = bswap32(s[i]);
      if (fwrite(buf, 4, c, f) != c) return false;
      s += c; n -= c;
    }
    return true;
  }
  bool put64(const void *src, size_t n) {
    const uint64_t *s = (const uint64_t *)src;
    uint64_t buf[2048];
    while (n) {
      const size_t c = std::min<size_t>(n, 2048);
      for (size_t i = 0; i < c; ++i) buf[i] = bswap64(s[i]);
      if (fwrite(buf, 8, c, f) != c) return false;
      s += c; n -= c;
    }
    return true;
  }
  bool get32(void *dst, size_t n) {
    uint32_t *d = (uint32_t *)dst;
    if (fread(d, 4, n, f) != n) return false;
    for (size_t i = 0; i < n; ++i) d[i] = bswap32(d[i]);
    return true;
  }
  bool get64(void *dst, size_t n) {
    uint64_t *d = (uint64_t *)dst;
    if (fread(d, 8, n, f) != n) return false;
    for (size_t i = 0; i < n; ++i) d[i] = bswap64(d[i]);
    return true;
  }
};

int open_error(const char *what, const std::string &path) {
  return fail(-90, std::string(what) + " path=" + path + ", iomsg=" + strerror(errno));
}

}  // namespace
}  // namespace tfx

using namespace tfx;

extern "C" int tfx_create_sensit_directory(const char *dir) {
  // create_directory (file_utils.F90:31-40): mkdir -p
  std::string d = dir ? dir : ".";
  for (size_t i = 1; i <= d.size(); ++i) {
    if (i == d.size() || d[i] == '/') {
      const std::string sub = d.substr(0, i);
      if (mkdir(sub.c_str(), 0777) != 0 && errno != EEXIST) return open_error("Error in creating the directory!", sub);
    }
  }
  return 0;
}

extern "C" int tfx_write_sensit_file(const tfx_sensit_rows *rows, const char *dir) {
  TFX_TRY(ensure_init());
  if (!rows) return fail(-82, "write_sensit_file: null handle");
  if (!rows->unit_weights)
    return fail(-91, "write_sensit_file: the rows carry problem / data weights; the file stores the unweighted kernel "
                     "(assemble with problem_weight = 1 and data_weight = 1, read_sensitivity_kernel applies them)");
  const tfx_sensit_params &P = rows->par;
  const int32_t N = P.nx * P.ny * P.nz, nmc = P.nmodel_components, ndc = P.ndata_components;
  const int64_t nseg = (int64_t)rows->ndata_loc * ndc * nmc;
  if ((int64_t)rows->seg_end.size() != nseg) return fail(-92, "write_sensit_file: the row set was already consumed");
  const std::string path = rank_file(dir, P.problem_type, rows->nbproc, rows->myrank);
  BeFile F;
  if (!F.open(path, "wb")) return open_error("Error in creating the sensitivity file!", path);
  const int32_t hdr[5] = {rows->ndata_loc, P.ndata, N, rows->myrank, rows->nbproc};   // :183
  bool ok = F.put32(hdr, 5);
  // stream the entries back in chunks of whole segments
  const int64_t chunk_cap = (int64_t)1 << 24;
  std::vector<int32_t> cols;
  std::vector<float> vals;
  cudaStream_t st = ctx().stream;
  int64_t s0 = 0;
  while (ok && s0 < nseg) {
    const int64_t e0 = s0 ? rows->seg_end[(size_t)s0 - 1] : 0;
    int64_t s1 = s0 + 1;
    while (s1 < nseg && rows->seg_end[(size_t)s1] - e0 <= chunk_cap) ++s1;
    const int64_t e1 = rows->seg_end[(size_t)s1 - 1];
    const size_t cnt = (size_t)(e1 - e0);
    cols.resize(std::max<size_t>(cnt, 1)); vals.resize(std::max<size_t>(cnt, 1));
    if (cnt) {
      TFX_CUDA(cudaMemcpyAsync(cols.data(), rows->R.idx.p + e0, cnt * 4, cudaMemcpyDeviceToHost, st));
      TFX_CUDA(cudaMemcpyAsync(vals.data(), rows->R.val.p + e0, cnt * 4, cudaMemcpyDeviceToHost, st));
      TFX_CUDA(cudaStreamSynchronize(st));
    }
    for (int64_t s = s0; ok && s < s1; ++s) {
      const int64_t b = (s ? rows->seg_end[(size_t)s - 1] : 0) - e0, e = rows->seg_end[(size_t)s] - e0;
      const int32_t k = (int32_t)(s % nmc), d = (int32_t)((s / nmc) % ndc);
      const int32_t idata = rows->data0 + (int32_t)(s / ((int64_t)nmc * ndc)) + 1;
      const int32_t nel = (int32_t)(e - b);
      const int32_t desc[4] = {idata, nel, k + 1, d + 1};                                 // :306
      ok = F.put32(desc, 4);
      if (ok && nel > 0) {
        for (int64_t i = b; i < e; ++i) cols[(size_t)i] = cols[(size_t)i] - k * N + 1;      // 1-based cell, :264
        ok = F.put32(cols.data() + b, (size_t)nel) && F.put32(vals.data() + b, (size_t)nel);   // :308
      }
    }
    s0 = s1;
  }
  if (!ok) return fail(-93, "Error in writing the sensitivity file! path=" + path);
  return 0;
}

extern "C" int tfx_write_sensit_metadata(const tfx_sensit_params *par, const char *dir, int32_t nbproc,
                                         int32_t depth_weighting_type, double comp_error, int64_t nnz_total,
                                         const int32_t *sensit_nnz) {
  const tfx_sensit_params &P = *par;
  const int32_t N = P.nx * P.ny * P.nz;
  {
    const std::string path = join(dir, std::string("sensit_") + suffix(P.problem_type) + "_meta.txt");
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return open_error("Error in creating the sensitivity metadata file!", path);
    fprintf(f, " %d %d %d %d\n", P.nx, P.ny, P.nz, P.ndata);                                // :366-370
    fprintf(f, " %d %d %d\n", nbproc, 4 /* MATRIX_PRECISION */, depth_weighting_type);
    fprintf(f, " %d %.17g\n", P.compression_type, comp_error);
    fprintf(f, " %d %d\n", P.nmodel_components, P.ndata_components);
    fprintf(f, " %lld\n", (long long)nnz_total);
    fclose(f);
  }
  if (sensit_nnz) {
    const std::string path = join(dir, std::string("sensit_") + suffix(P.problem_type) + "_nnz");
    BeFile F;
    if (!F.open(path, "wb")) return open_error("Error in creating the sensit_nnz file!", path);
    if (!F.put32(&N, 1) || !F.put32(sensit_nnz, (size_t)N)) return fail(-93, "Error in writing the file! path=" + path);   // :388-389
  }
  return 0;
}

extern "C" int tfx_read_sensitivity_metadata(const tfx_sensit_params *par, const char *dir, int32_t depth_weighting_type,
                                             int32_t *nbproc_sensit, double *comp_error, int64_t *nnz_total) {
  const tfx_sensit_params &P = *par;
  const std::string path = join(dir, std::string("sensit_") + suffix(P.problem_type) + "_meta.txt");
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return open_error("Error in opening the sensitivity metadata file!", path);
  int nx, ny, nz, nd, nb, prec, wt, ct, nmc, ndc;
  double ce = 0.0;
  long long nt = 0;
  // list-directed input: blanks, commas and line ends separate the items; gfortran may print D exponents
  char buf[64];
  auto next_num = [&](double *out) -> bool {
    if (fscanf(f, " %63[^ ,\n\r\t]%*[ ,\n\r\t]", buf) < 1) return false;
    for (char *c = buf; *c; ++c) if (*c == 'D' || *c == 'd') *c = 'e';
    char *end = nullptr;
    *out = strtod(buf, &end);
    return end != buf;
  };
  double v[12];
  int got = 0;
  for (; got < 12; ++got) if (!next_num(&v[got])) break;
  fclose(f);
  if (got < 11) return fail(-94, "Error in reading the sensitivity metadata file! path=" + path);
  nx = (int)v[0]; ny = (int)v[1]; nz = (int)v[2]; nd = (int)v[3]; nb = (int)v[4]; prec = (int)v[5]; wt = (int)v[6];
  ct = (int)v[7]; ce = v[8]; nmc = (int)v[9]; ndc = (int)v[10];
  nt = got >= 12 ? (long long)v[11] : 0;
  if (nx != P.nx || ny != P.ny || nz != P.nz || nd != P.ndata || wt != depth_weighting_type ||
      nmc != P.nmodel_components || ndc != P.ndata_components)
    return fail(-95, "Sensitivity metadata file info does not match the Parfile!");          // :1014-1018
  if (ct != P.compression_type) return fail(-95, "Compression type is inconsistent!");        // :1020-1022
  if (prec != 4) return fail(-95, "Matrix precision is not consistent!");                     // :1025-1027
  if (nbproc_sensit) *nbproc_sensit = nb;
  if (comp_error) *comp_error = ce;
  if (nnz_total) *nnz_total = nt;
  return 0;
}

extern "C" int tfx_read_sensit_nnz(const tfx_sensit_params *par, const char *dir, int32_t *sensit_nnz) {
  const tfx_sensit_params &P = *par;
  const int32_t N = P.nx * P.ny * P.nz;
  const std::string path = join(dir, std::string("sensit_") + suffix(P.problem_type) + "_nnz");
  BeFile F;
  if (!F.open(path, "rb")) return open_error("Error in opening the sensitivity file!", path);
  int32_t n = 0;
  if (!F.get32(&n, 1) || n != N) return fail(-95, "Wrong file header in calculate_new_partitioning!");   // :556-558
  if (!F.get32(sensit_nnz, (size_t)N)) return fail(-94, "Error in reading the file! path=" + path);
  return 0;
}

extern "C" int tfx_write_depth_weight(const tfx_sensit_params *par, const char *dir, const double *column_weight_full) {
  const tfx_sensit_params &P = *par;
  const int32_t N = P.nx * P.ny * P.nz;
  TFX_TRY(tfx_create_sensit_directory(dir));
  const std::string path = join(dir, std::string("sensit_") + suffix(P.problem_type) + "_weight");
  BeFile F;
  if (!F.open(path, "wb")) return open_error("Error in creating the depth weight file!", path);
  if (!F.put32(&N, 1) || !F.put64(column_weight_full, (size_t)N)) return fail(-93, "Error in writing the file! path=" + path);   // :457-458
  return 0;
}

extern "C" int tfx_read_depth_weight(const tfx_sensit_params *par, const char *dir, double *column_weight_full) {
  const tfx_sensit_params &P = *par;
  const int32_t N = P.nx * P.ny * P.nz;
  const std::string path = join(dir, std::string("sensit_") + suffix(P.problem_type) + "_weight");
  BeFile F;
  if (!F.open(path, "rb")) return open_error("Error in opening the depth weight file!", path);
  int32_t n = 0;
  if (!F.get32(&n, 1)) return fail(-94, "Error in reading the file! path=" + path);
  if (!F.get64(column_weight_full, (size_t)std::min(n, N)) || n != N)
    return fail(-95, "Depth weight file header does not match the Parfile!");                  // :955-957
  return 0;
}

static int read_kernel_core(const tfx_sensit_params *par, const char *dir, const double *data_weight,
                            int32_t depth_weighting_type, int32_t problem_slot, int32_t myrank, int32_t nbproc,
                            const int32_t *nelements_at_cpu, RowTriplets &R, int32_t *nl_out, int32_t *ncolumns_out) {
  TFX_TRY(ensure_init());
  if (!par) return fail(-82, "read_sensitivity_kernel: null handle");
  if (problem_slot != 1 && problem_slot != 2) return fail(-85, "read_sensitivity_kernel: problem_slot must be 1 or 2");
  if (nbproc < 1 || myrank < 0 || myrank >= nbproc) return fail(-83, "read_sensitivity_kernel: wrong rank");
  const tfx_sensit_params &P = *par;
  const int32_t N = P.nx * P.ny * P.nz, nmc = P.nmodel_components, ndc = P.ndata_components;
  const int32_t nel_compressed = (P.compression_type > 0) ? (int32_t)(P.compression_rate * (double)N) : N;
  int32_t nsmaller = 0, total = 0;
  for (int32_t r = 0; r < nbproc; ++r) {
    if (r < myrank) nsmaller += nelements_at_cpu[r];
    total += nelements_at_cpu[r];
  }
  if (total != N) return fail(-87, "read_sensitivity_kernel: nelements_at_cpu does not sum to nx*ny*nz");
  const int32_t nel_loc = nelements_at_cpu[myrank];
  const int32_t param_shift = (problem_slot - 1) * nel_loc * nmc;                               // :685-686
  int32_t nbproc_sensit = 0;
  TFX_TRY(tfx_read_sensitivity_metadata(par, dir, depth_weighting_type, &nbproc_sensit, nullptr, nullptr));

  std::vector<int32_t> h_idx, h_row, cols((size_t)std::max(nel_compressed, 1));
  std::vector<float> h_val, vals((size_t)std::max(nel_compressed, 1));
  int32_t idata_glob = 0;
  for (int32_t rank = 0; rank < nbproc_sensit; ++rank) {
    const std::string path = rank_file(dir, P.problem_type, nbproc_sensit, rank);
    BeFile F;
    if (!F.open(path, "rb")) return open_error("Error in opening the sensitivity file!", path);
    int32_t hdr[5];
    if (!F.get32(hdr, 5)) return fail(-94, "Error in reading the sensitivity file! path=" + path);
    if (hdr[1] != P.ndata || hdr[2] != N || hdr[3] != rank || hdr[4] != nbproc_sensit)
      return fail(-95, "Wrong file header in read_sensitivity_kernel!");                        // :749-752
    for (int32_t i = 0; i < hdr[0]; ++i) {
      ++idata_glob;
      for (int32_t d = 1; d <= ndc; ++d) {
        const int32_t row = (idata_glob - 1) * ndc + (d - 1);
        // combined_weight = real(problem_weight * data_weight(d, idata_glob), 4) (:837)
        const float wgt = (float)(P.problem_weight * data_weight[(size_t)(idata_glob - 1) * ndc + (d - 1)]);
        for (int32_t k = 1; k <= nmc; ++k) {
          int32_t desc[4];
          if (!F.get32(desc, 4)) return fail(-94, "Error in reading the sensitivity file! path=" + path);
          if (desc[0] != idata_glob) return fail(-95, "Wrong data index in read_sensitivity_kernel!");            // :772-774
          if (desc[1] > nel_compressed || desc[1] < 0)
            return fail(-95, "Wrong number of elements in read_sensitivity_kernel!");                              // :777-779
          if (desc[2] != k) return fail(-95, "Wrong model component index in read_sensitivity_kernel!");          // :782-784
          if (desc[3] != d) return fail(-95, "Wrong data component index in read_sensitivity_kernel!");           // :787-789
          const int32_t nel = desc[1];
          if (nel > 0 && (!F.get32(cols.data(), (size_t)nel) || !F.get32(vals.data(), (size_t)nel)))
            return fail(-94, "Error in reading the sensitivity file! path=" + path);
          // this rank's piece: cells nsmaller < p <= nsmaller + nelements (:796-803; columns ascend)
          const int32_t *b = std::upper_bound(cols.data(), cols.data() + nel, nsmaller);
          const int32_t *e = std::upper_bound(cols.data(), cols.data() + nel, nsmaller + nel_loc);
          const int32_t index_shift = param_shift + (k - 1) * nel_loc - nsmaller;                                   // :834
          for (const int32_t *c = b; c < e; ++c) {
            h_idx.push_back(*c + index_shift - 1);                                                                  // 0-based
            h_val.push_back(vals[(size_t)(c - cols.data())] * wgt);                                                 // real(4) product (:842)
            h_row.push_back(row);
          }
        }
      }
    }
  }
  if (idata_glob != P.ndata) return fail(-95, "Wrong number of data rows in read_sensitivity_kernel!");

  const size_t nnz = h_idx.size();
  cudaStream_t st = ctx().stream;
  TFX_TRY(R.idx.alloc(std::max<size_t>(nnz, 1))); TFX_TRY(R.val.alloc(std::max<size_t>(nnz, 1)));
  TFX_TRY(R.rowid.alloc(std::max<size_t>(nnz, 1)));
  if (nnz) {
    TFX_CUDA(cudaMemcpyAsync(R.idx.p, h_idx.data(), nnz * 4, cudaMemcpyHostToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(R.val.p, h_val.data(), nnz * 4, cudaMemcpyHostToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(R.rowid.p, h_row.data(), nnz * 4, cudaMemcpyHostToDevice, st));
    TFX_CUDA(cudaStreamSynchronize(st));
  }
  R.nnz = (int64_t)nnz;
  *nl_out = P.ndata * ndc;
  *ncolumns_out = 2 * nmc * nel_loc;
  return 0;
}

extern "C" int tfx_read_sensitivity_kernel(tfx_matrix **out, const tfx_sensit_params *par, const char *dir,
                                           const double *data_weight, int32_t depth_weighting_type,
                                           int32_t problem_slot, int32_t myrank, int32_t nbproc,
                                           const int32_t *nelements_at_cpu, int64_t *nnz_local) {
  if (!out) return fail(-82, "read_sensitivity_kernel: null handle");
  RowTriplets R;
  int32_t nl = 0, ncolumns = 0;
  TFX_TRY(read_kernel_core(par, dir, data_weight, depth_weighting_type, problem_slot, myrank, nbproc, nelements_at_cpu, R,
                           &nl, &ncolumns));
  if (nnz_local) *nnz_local = R.nnz;
  tfx_matrix *h = new tfx_matrix();
  int rc = matrix_from_triplets(h->m, nl, ncolumns, R);
  if (rc) { delete h; return rc; }
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  *out = h;
  return 0;
}

// The reference's calling convention: rows are appended to the matrix under construction (one call per problem,
// problem_joint_gravmag.F90:241-248); sparse_matrix finalize() builds the device representations.
extern "C" int tfx_read_sensitivity_kernel_into(tfx_matrix *matrix_sensit, const tfx_sensit_params *par, const char *dir,
                                                const double *data_weight, int32_t depth_weighting_type,
                                                int32_t problem_slot, int32_t myrank, int32_t nbproc,
                                                const int32_t *nelements_at_cpu, int64_t *nnz_local) {
  if (!matrix_sensit) return fail(-82, "read_sensitivity_kernel_into: null handle");
  RowTriplets R;
  int32_t nl = 0, ncolumns = 0;
  TFX_TRY(read_kernel_core(par, dir, data_weight, depth_weighting_type, problem_slot, myrank, nbproc, nelements_at_cpu, R,
                           &nl, &ncolumns));
  if (nnz_local) *nnz_local = R.nnz;
  if (g_opt_sensit_row_blocks) return matrix_append_block(matrix_sensit->m, R, nl, ncolumns);
  return matrix_append_triplets(matrix_sensit->m, R, nl, ncolumns);
}
