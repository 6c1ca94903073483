#include "smc_internal.h"
extern "C" int smc_nuts_run(smc_model_t m, const smc_sampler_config* cfg, const double* init, int64_t C, smc_run_output* out) {
  (void)m; (void)cfg; (void)init; (void)C; (void)out;
  smc_set_error("smc_nuts_run: not built yet");
  return SMC_ERR_INVALID;
}
