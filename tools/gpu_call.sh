tools/potrf_bench
