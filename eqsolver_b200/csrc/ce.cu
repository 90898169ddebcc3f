// ce.cu -- TEMPORARY stub while the CrossEntropy kernels are being written (replaced in the next commit).
#include "common.cuh"
using namespace eqs;
struct eqs_ctx { Ctx c; };
extern "C" {
int eqs_ce_validate(const eqs_ce_params *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_solve(eqs_ctx *, const eqs_ce_params *, const double *, double *, eqs_report *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_create(eqs_ctx *, const eqs_ce_params *, const double *, eqs_ce **) { return EQS_ERR_EXTERNAL; }
int eqs_ce_destroy(eqs_ce *) { return EQS_OK; }
int eqs_ce_step(eqs_ce *, int64_t, const double *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_sync(eqs_ce *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_last_step_ms(eqs_ce *, double *, int64_t *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_read_state(eqs_ce *, double *, double *, int64_t *, int32_t *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_read_elites(eqs_ce *, int64_t *, int64_t *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_read_costs(eqs_ce *, double *) { return EQS_ERR_EXTERNAL; }
int eqs_ce_local_range(eqs_ce *, int64_t *, int64_t *) { return EQS_ERR_EXTERNAL; }
}
