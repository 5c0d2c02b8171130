// temporary: K2 entry points until resect.cu lands
#include "context.cuh"
extern "C" {
int hulo_score_resection(hulo_gpu *, const double *, size_t, const double *, const double *, size_t, const double *, double, float *, int32_t *, float *, int32_t *) { hulo::set_error("not built yet"); return HULO_ERR_ARG; }
int hulo_resection_residuals(hulo_gpu *, const double *, size_t, const double *, const double *, size_t, const double *, float *) { hulo::set_error("not built yet"); return HULO_ERR_ARG; }
int hulo_p3p(hulo_gpu *, const uint32_t *, size_t, const double *, const double *, size_t, const double *, double *, int32_t *) { hulo::set_error("not built yet"); return HULO_ERR_ARG; }
int hulo_resect_acransac(hulo_gpu *, const double *, const double *, size_t, const double *, size_t, uint64_t, double *, int32_t *, size_t *, double *, int *) { hulo::set_error("not built yet"); return HULO_ERR_ARG; }
}
