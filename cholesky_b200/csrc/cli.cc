// `cholesky` command line: the reference's flags (mmat.rg:1072-1093) over the C ABI.
//   -i matrix.mtx -s separators.txt -c clusters.txt [-b rhs.mtx -o solution] [-m factor.mtx]
//   [-p permuted.mtx] [-d debug_dir] [--iterations N] [--gpu D | --gpus N] [--binary-factor factor.bin]
// --gpus N (2, 4 or 8) runs the factorization and the solve subtree-partitioned over CUDA devices 0 .. N-1 from this
// one process (the counterpart of the reference's -ll:gpu processor selection), --devices d0,d1,.. over an explicit
// list (a device may repeat: several ranks then share it); --gpu D picks one device.
//   cholesky --convert factor.bin factor.mtx      (binary dump -> the reference's text format, no GPU)
// Progress lines follow the reference's stdout (mmat.rg:1095-1121, 1228, 1357).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/cholesky.h"

int main(int argc, char **argv) {
  const char *mat = "", *sep = "", *clu = "", *bfile = "", *sol = "", *fac = "", *perm = "", *dbg = "";
  int iterations = 1, gpu = 0, gpus = 1;
  const char *devlist = "";
  const char *binfac = "";
  if (argc == 4 && !strcmp(argv[1], "--convert")) {
    int rc = chol_factor_binary_to_mtx(argv[2], argv[3], 0);
    if (rc) printf("cannot convert %s: error %d\n", argv[2], rc);
    return rc ? 1 : 0;
  }
  for (int i = 0; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "-i")) mat = argv[i + 1];
    else if (!strcmp(argv[i], "-s")) sep = argv[i + 1];
    else if (!strcmp(argv[i], "-c")) clu = argv[i + 1];
    else if (!strcmp(argv[i], "-m")) fac = argv[i + 1];
    else if (!strcmp(argv[i], "-p")) perm = argv[i + 1];
    else if (!strcmp(argv[i], "-o")) sol = argv[i + 1];
    else if (!strcmp(argv[i], "-b")) bfile = argv[i + 1];
    else if (!strcmp(argv[i], "-d")) dbg = argv[i + 1];  // debug_path, debug = true (mmat.rg:1086-1090)
    else if (!strcmp(argv[i], "--iterations")) iterations = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--gpu")) gpu = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--gpus")) gpus = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "--devices")) devlist = argv[i + 1];
    else if (!strcmp(argv[i], "--binary-factor")) binfac = argv[i + 1];
  }
  printf("Iterations: %d\n", iterations);
  chol_t *c = nullptr;
  std::vector<int> devices;
  for (int d = 0; d < gpus; d++) devices.push_back(gpus > 1 ? d : gpu);
  if (*devlist) {
    devices.clear();
    for (const char *q = devlist; *q;) {
      devices.push_back(atoi(q));
      while (*q && *q != ',') q++;
      if (*q == ',') q++;
    }
  }
  if (chol_create(devices.data(), (int)devices.size(), &c)) {
    printf("the factorization runs on 1, 2, 4 or 8 GPUs\n");
    return 1;
  }
  if (devices.size() > 1) printf("GPUs: %d\n", (int)devices.size());
  if (chol_load(c, mat, sep, clu)) {
    printf("%s\n", chol_last_error(c));
    return 1;
  }
  printf("M: %d N: %d nz: %lld\n", chol_n(c), chol_n(c), (long long)chol_nz(c));
  if (chol_analyze(c, *dbg ? 1 : 0)) {
    printf("%s\n", chol_last_error(c));
    return 1;
  }
  printf("levels: %d\nseparators: %d\nMax Interval Size: %d\n", chol_levels(c), chol_num_separators(c), chol_max_int_size(c));
  printf("Blocks ispace: %lld\nClusters ispace: %lld\n", (long long)chol_num_blocks(c), (long long)chol_num_clusters0(c));
  if (*perm) {
    if (chol_assemble(c) || chol_write_factor(c, perm, 0)) {
      printf("%s\n", chol_last_error(c));
      return 1;
    }
    printf("saving matrix to: %s\n\n", perm);
  }
  chol_stats_t st;
  if (*dbg) {
    // the debug run: log lines on stdout (redirect them to the file verify.debug_factor reads), one
    // snapshot per fused task under the debug path, then the regular outputs from the same factor
    if (chol_write_debug_log(c, nullptr) || chol_factor_debug(c, dbg, 0, 1)) {
      printf("%s\n", chol_last_error(c));
      return 1;
    }
    st.seconds_best = 0, st.flops = 0, st.kernel_launches = 0;
  } else if (chol_factor(c, iterations, 0, &st)) {
    printf("%s\n", chol_last_error(c));
    return 1;
  }
  if (!*dbg)
    printf("Done factoring: %.6f s (best of %d), %.3f GFLOP/s, %lld kernels per iteration\n", st.seconds_best, iterations,
         st.flops / st.seconds_best * 1e-9, (long long)st.kernel_launches);
  if (*fac) {
    printf("saving matrix to: %s\n\n", fac);
    if (chol_write_factor(c, fac, 0)) {
      printf("%s\n", chol_last_error(c));
      return 1;
    }
  }
  if (*binfac) {
    printf("saving matrix to: %s\n\n", binfac);
    if (chol_write_factor_binary(c, binfac)) {
      printf("%s\n", chol_last_error(c));
      return 1;
    }
  }
  if (*bfile) {
    int n = chol_n(c);
    std::vector<double> b(n), x(n);
    if (chol_read_vector(bfile, n, b.data())) {
      printf("cannot read %s\n", bfile);
      return 1;
    }
    if (chol_solve(c, b.data(), x.data())) {
      printf("%s\n", chol_last_error(c));
      return 1;
    }
    printf("Done solve.\n");
    if (*sol) {
      printf("Saving solution to: %s\n", sol);
      chol_write_solution(sol, n, x.data());
    }
  }
  chol_destroy(c);
  return 0;
}
