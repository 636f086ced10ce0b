// libbppp.so, latency build of the 4-lane ladder kernels: engine_var.cu compiled with the call boundary at the point operation
// (see the note at the top of engine_var.cu).  Selected for small (sub-)batches by engine_var.cu:var_lanes_for.
#define BPPP_VAR_LAT 1
#include "engine_var.cu"
