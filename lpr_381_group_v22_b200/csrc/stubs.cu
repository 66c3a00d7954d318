// stubs.cu -- TEMPORARY: entry points not implemented yet return LPR_E_STATE.
#include "common.cuh"
extern "C" {
int lpr_rev_refactor(lpr_rev* h) { return lpr::fail(LPR_E_STATE, "lpr_rev_refactor: not implemented yet"); }
int lpr_tab_round4(lpr_tab* h) { return lpr::fail(LPR_E_STATE, "lpr_tab_round4: not implemented yet"); }
int lpr_tab_bb_node_solve(lpr_tab* h, int64_t max_pivots, int* status, int64_t* n_pivots, int* pivot_log, int64_t log_cap) { return lpr::fail(LPR_E_STATE, "lpr_tab_bb_node_solve: not implemented yet"); }
int lpr_tab_bb_add_constraint(lpr_tab* parent, int n_vars, int var, double bound, int type, lpr_tab** child) { return lpr::fail(LPR_E_STATE, "lpr_tab_bb_add_constraint: not implemented yet"); }
int lpr_tab_bb_branch_var(lpr_tab* h, int n_vars, int* var, double* value, double* x /* n_vars, may be NULL */) { return lpr::fail(LPR_E_STATE, "lpr_tab_bb_branch_var: not implemented yet"); }
int lpr_bb_solve(int device, int rows, int cols, const double* final_tableau, int n_vars, int enable_pruning, int64_t max_nodes, double* x, double* z, int* has_solution, int64_t* nodes, int64_t* pivots, int* node_log, double* node_z, int64_t node_log_cap, int* status) { return lpr::fail(LPR_E_STATE, "lpr_bb_solve: not implemented yet"); }
int lpr_bb_create(int device, int rows, int cols, const double* root_tableau, int n_vars, int enable_pruning, lpr_bb** out) { return lpr::fail(LPR_E_STATE, "lpr_bb_create: not implemented yet"); }
int lpr_bb_destroy(lpr_bb* h) { return lpr::fail(LPR_E_STATE, "lpr_bb_destroy: not implemented yet"); }
int lpr_bb_open_count(lpr_bb* h, int64_t* n) { return lpr::fail(LPR_E_STATE, "lpr_bb_open_count: not implemented yet"); }
int lpr_bb_run(lpr_bb* h, int64_t max_nodes, int64_t* processed, int64_t* pivots) { return lpr::fail(LPR_E_STATE, "lpr_bb_run: not implemented yet"); }
int lpr_bb_get_incumbent(lpr_bb* h, int* has, double* z, double* x, int* key, int* key_len) { return lpr::fail(LPR_E_STATE, "lpr_bb_get_incumbent: not implemented yet"); }
int lpr_bb_set_incumbent(lpr_bb* h, double z, const double* x, const int* key, int key_len) { return lpr::fail(LPR_E_STATE, "lpr_bb_set_incumbent: not implemented yet"); }
int lpr_bb_export_nodes(lpr_bb* h, int max_nodes, void* buf, int64_t buf_cap, int64_t* bytes, int* n_exported) { return lpr::fail(LPR_E_STATE, "lpr_bb_export_nodes: not implemented yet"); }
int lpr_bb_import_nodes(lpr_bb* h, const void* buf, int64_t bytes) { return lpr::fail(LPR_E_STATE, "lpr_bb_import_nodes: not implemented yet"); }
int lpr_knap_dp(int device, int capacity, int n, const int* weights, const int* values, double* best, uint8_t* chosen) { return lpr::fail(LPR_E_STATE, "lpr_knap_dp: not implemented yet"); }
int lpr_knap_create(int device, double capacity, int n, const double* weights, const double* values, lpr_knap** out) { return lpr::fail(LPR_E_STATE, "lpr_knap_create: not implemented yet"); }
int lpr_knap_destroy(lpr_knap* h) { return lpr::fail(LPR_E_STATE, "lpr_knap_destroy: not implemented yet"); }
int lpr_knap_run(lpr_knap* h, int64_t max_nodes, int64_t* processed, int* status) { return lpr::fail(LPR_E_STATE, "lpr_knap_run: not implemented yet"); }
int lpr_knap_open_count(lpr_knap* h, int64_t* n) { return lpr::fail(LPR_E_STATE, "lpr_knap_open_count: not implemented yet"); }
int lpr_knap_get_incumbent(lpr_knap* h, double* best, uint8_t* chosen /* n, original ids */, uint64_t* key /* key_words */, int* key_bits) { return lpr::fail(LPR_E_STATE, "lpr_knap_get_incumbent: not implemented yet"); }
int lpr_knap_set_incumbent(lpr_knap* h, double best, const uint8_t* chosen, const uint64_t* key, int key_bits) { return lpr::fail(LPR_E_STATE, "lpr_knap_set_incumbent: not implemented yet"); }
int lpr_knap_export_nodes(lpr_knap* h, int max_nodes, void* buf, int64_t buf_cap, int64_t* bytes, int* n_exported) { return lpr::fail(LPR_E_STATE, "lpr_knap_export_nodes: not implemented yet"); }
int lpr_knap_import_nodes(lpr_knap* h, const void* buf, int64_t bytes) { return lpr::fail(LPR_E_STATE, "lpr_knap_import_nodes: not implemented yet"); }
int lpr_knap_solve(int device, double capacity, int n, const double* weights, const double* values, int64_t max_nodes, double* best, uint8_t* chosen, int64_t* nodes, int* status) { return lpr::fail(LPR_E_STATE, "lpr_knap_solve: not implemented yet"); }
}
