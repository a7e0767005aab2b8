// Empty stand-in: ivf_pq_fp_8bit.cuh includes this third-party / library header but fp_8bit itself uses nothing from it
// (RAFT, rmm and the cuvs public headers' dependencies are not available offline).  See ../Makefile.
