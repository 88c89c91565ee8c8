// ORACLE (test infrastructure).  Link stubs for the two CUDA launchers the reference's chamfer_distance.cpp declares
// (chamfer_distance.cpp:3-24) and defines in chamfer_distance.cu.  oracle/build_ref.py compiles only the reference's
// CPU entry points (forward / backward); these stubs make the module link and fail loudly if the CUDA names are called.
int ChamferDistanceKernelLauncher(const int, const int, const float*, const int, const float*, float*, int*, float*,
                                  int*) {
  return -1;
}
int ChamferDistanceGradKernelLauncher(const int, const int, const float*, const int, const float*, const float*,
                                      const int*, const float*, const int*, float*, float*) {
  return -1;
}
