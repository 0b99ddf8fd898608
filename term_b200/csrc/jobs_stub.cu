// Jobs not implemented yet report TG_ERR_UNSUPPORTED on the aggregate (never a silent CPU fallback).
#include "engine.hpp"
namespace tg {
void exec_distinct_job(Engine&, Table&, Plan&, int) { throw Error(TG_ERR_UNSUPPORTED, "distinct/unique job not implemented yet"); }
void exec_fk_job(Engine&, Plan&, int) { throw Error(TG_ERR_UNSUPPORTED, "foreign key job not implemented yet"); }
void exec_kll_job(Engine&, Table&, Plan&, int) { throw Error(TG_ERR_UNSUPPORTED, "KLL job not implemented yet"); }
void exec_grouped_job(Engine&, Table&, Plan&, int) { throw Error(TG_ERR_UNSUPPORTED, "grouped job not implemented yet"); }
void exec_spearman_job(Engine&, Table&, Plan&, int) { throw Error(TG_ERR_UNSUPPORTED, "spearman job not implemented yet"); }
}  // namespace tg
