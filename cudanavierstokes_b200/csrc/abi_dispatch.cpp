// libcudns: the public C symbols of include/cudns.h that take a solver handle.  The device side exists twice -- api.cu and the kernels
// compiled with real = double (cudns64_*) and with real = float (cudns32_*; `myprec` of the reference, src/globals.h:5-6, as a
// run-time choice: cudns_params.precision) -- and every solver object starts with an int that says which copy it belongs to.
// Generated from the header's prototypes (tools/gen_abi_dispatch.py): one forwarding function per entry point, nothing else.
#include "../../include/cudns.h"

extern "C" {
int cudns64_create(const cudns_params *p, const double *x, const double *xp, const double *xpp, cudns_handle *out);
int cudns32_create(const cudns_params *p, const double *x, const double *xp, const double *xpp, cudns_handle *out);
int cudns_create(const cudns_params *p, const double *x, const double *xp, const double *xpp, cudns_handle *out) { return (p && p->precision == 1) ? cudns32_create(p, x, xp, xpp, out) : cudns64_create(p, x, xp, xpp, out); }
int cudns64_destroy(cudns_handle h);
int cudns32_destroy(cudns_handle h);
int cudns_destroy(cudns_handle h) { return (h && *(const int *)h == 1) ? cudns32_destroy(h) : cudns64_destroy(h); }
int cudns64_memory_report(cudns_handle h, size_t *solver_bytes, size_t *free_bytes, size_t *total_bytes);
int cudns32_memory_report(cudns_handle h, size_t *solver_bytes, size_t *free_bytes, size_t *total_bytes);
int cudns_memory_report(cudns_handle h, size_t *solver_bytes, size_t *free_bytes, size_t *total_bytes) { return (h && *(const int *)h == 1) ? cudns32_memory_report(h, solver_bytes, free_bytes, total_bytes) : cudns64_memory_report(h, solver_bytes, free_bytes, total_bytes); }
int cudns64_set_state(cudns_handle h, const double *r, const double *u, const double *v, const double *w, const double *e);
int cudns32_set_state(cudns_handle h, const double *r, const double *u, const double *v, const double *w, const double *e);
int cudns_set_state(cudns_handle h, const double *r, const double *u, const double *v, const double *w, const double *e) { return (h && *(const int *)h == 1) ? cudns32_set_state(h, r, u, v, w, e) : cudns64_set_state(h, r, u, v, w, e); }
int cudns64_get_state(cudns_handle h, double *r, double *u, double *v, double *w, double *e);
int cudns32_get_state(cudns_handle h, double *r, double *u, double *v, double *w, double *e);
int cudns_get_state(cudns_handle h, double *r, double *u, double *v, double *w, double *e) { return (h && *(const int *)h == 1) ? cudns32_get_state(h, r, u, v, w, e) : cudns64_get_state(h, r, u, v, w, e); }
int cudns64_set_state_device(cudns_handle h, const double *d_r, const double *d_u, const double *d_v, const double *d_w, const double *d_e);
int cudns32_set_state_device(cudns_handle h, const double *d_r, const double *d_u, const double *d_v, const double *d_w, const double *d_e);
int cudns_set_state_device(cudns_handle h, const double *d_r, const double *d_u, const double *d_v, const double *d_w, const double *d_e) { return (h && *(const int *)h == 1) ? cudns32_set_state_device(h, d_r, d_u, d_v, d_w, d_e) : cudns64_set_state_device(h, d_r, d_u, d_v, d_w, d_e); }
int cudns64_get_state_device(cudns_handle h, double *d_r, double *d_u, double *d_v, double *d_w, double *d_e);
int cudns32_get_state_device(cudns_handle h, double *d_r, double *d_u, double *d_v, double *d_w, double *d_e);
int cudns_get_state_device(cudns_handle h, double *d_r, double *d_u, double *d_v, double *d_w, double *d_e) { return (h && *(const int *)h == 1) ? cudns32_get_state_device(h, d_r, d_u, d_v, d_w, d_e) : cudns64_get_state_device(h, d_r, d_u, d_v, d_w, d_e); }
int cudns64_set_sponge(cudns_handle h, const double *sigma_x, const double *sigma_z, const double *ref5);
int cudns32_set_sponge(cudns_handle h, const double *sigma_x, const double *sigma_z, const double *ref5);
int cudns_set_sponge(cudns_handle h, const double *sigma_x, const double *sigma_z, const double *ref5) { return (h && *(const int *)h == 1) ? cudns32_set_sponge(h, sigma_x, sigma_z, ref5) : cudns64_set_sponge(h, sigma_x, sigma_z, ref5); }
int cudns64_advance(cudns_handle h, int nsteps, double *time, double *par1, double *par2);
int cudns32_advance(cudns_handle h, int nsteps, double *time, double *par1, double *par2);
int cudns_advance(cudns_handle h, int nsteps, double *time, double *par1, double *par2) { return (h && *(const int *)h == 1) ? cudns32_advance(h, nsteps, time, par1, par2) : cudns64_advance(h, nsteps, time, par1, par2); }
int cudns64_calc_rhs(cudns_handle h, double *rhs_r, double *rhs_u, double *rhs_v, double *rhs_w, double *rhs_e);
int cudns32_calc_rhs(cudns_handle h, double *rhs_r, double *rhs_u, double *rhs_v, double *rhs_w, double *rhs_e);
int cudns_calc_rhs(cudns_handle h, double *rhs_r, double *rhs_u, double *rhs_v, double *rhs_w, double *rhs_e) { return (h && *(const int *)h == 1) ? cudns32_calc_rhs(h, rhs_r, rhs_u, rhs_v, rhs_w, rhs_e) : cudns64_calc_rhs(h, rhs_r, rhs_u, rhs_v, rhs_w, rhs_e); }
int cudns64_calc_dt(cudns_handle h, double *dt);
int cudns32_calc_dt(cudns_handle h, double *dt);
int cudns_calc_dt(cudns_handle h, double *dt) { return (h && *(const int *)h == 1) ? cudns32_calc_dt(h, dt) : cudns64_calc_dt(h, dt); }
int cudns64_calc_bulk(cudns_handle h, double *par1, double *par2);
int cudns32_calc_bulk(cudns_handle h, double *par1, double *par2);
int cudns_calc_bulk(cudns_handle h, double *par1, double *par2) { return (h && *(const int *)h == 1) ? cudns32_calc_bulk(h, par1, par2) : cudns64_calc_bulk(h, par1, par2); }
int cudns64_calc_enstrophy(cudns_handle h, double *enstrophy);
int cudns32_calc_enstrophy(cudns_handle h, double *enstrophy);
int cudns_calc_enstrophy(cudns_handle h, double *enstrophy) { return (h && *(const int *)h == 1) ? cudns32_calc_enstrophy(h, enstrophy) : cudns64_calc_enstrophy(h, enstrophy); }
int cudns64_get_scalars(cudns_handle h, double *dt, double *dpdz, double *time);
int cudns32_get_scalars(cudns_handle h, double *dt, double *dpdz, double *time);
int cudns_get_scalars(cudns_handle h, double *dt, double *dpdz, double *time) { return (h && *(const int *)h == 1) ? cudns32_get_scalars(h, dt, dpdz, time) : cudns64_get_scalars(h, dt, dpdz, time); }
int cudns64_set_dt(cudns_handle h, double dt, int fixed);
int cudns32_set_dt(cudns_handle h, double dt, int fixed);
int cudns_set_dt(cudns_handle h, double dt, int fixed) { return (h && *(const int *)h == 1) ? cudns32_set_dt(h, dt, fixed) : cudns64_set_dt(h, dt, fixed); }
int cudns64_halo_local_info(cudns_handle h, cudns_peer_info *mine);
int cudns32_halo_local_info(cudns_handle h, cudns_peer_info *mine);
int cudns_halo_local_info(cudns_handle h, cudns_peer_info *mine) { return (h && *(const int *)h == 1) ? cudns32_halo_local_info(h, mine) : cudns64_halo_local_info(h, mine); }
int cudns64_halo_connect(cudns_handle h, const cudns_peer_info *lower, const cudns_peer_info *upper);
int cudns32_halo_connect(cudns_handle h, const cudns_peer_info *lower, const cudns_peer_info *upper);
int cudns_halo_connect(cudns_handle h, const cudns_peer_info *lower, const cudns_peer_info *upper) { return (h && *(const int *)h == 1) ? cudns32_halo_connect(h, lower, upper) : cudns64_halo_connect(h, lower, upper); }
int cudns64_halo_buffers(cudns_handle h, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, size_t *bytes_each);
int cudns32_halo_buffers(cudns_handle h, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, size_t *bytes_each);
int cudns_halo_buffers(cudns_handle h, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, size_t *bytes_each) { return (h && *(const int *)h == 1) ? cudns32_halo_buffers(h, send_lo, send_hi, recv_lo, recv_hi, bytes_each) : cudns64_halo_buffers(h, send_lo, send_hi, recv_lo, recv_hi, bytes_each); }
int cudns64_set_allreduce(cudns_handle h, cudns_allreduce_fn fn, void *user);
int cudns32_set_allreduce(cudns_handle h, cudns_allreduce_fn fn, void *user);
int cudns_set_allreduce(cudns_handle h, cudns_allreduce_fn fn, void *user) { return (h && *(const int *)h == 1) ? cudns32_set_allreduce(h, fn, user) : cudns64_set_allreduce(h, fn, user); }
int cudns64_set_exchange(cudns_handle h, cudns_exchange_fn fn, void *user);
int cudns32_set_exchange(cudns_handle h, cudns_exchange_fn fn, void *user);
int cudns_set_exchange(cudns_handle h, cudns_exchange_fn fn, void *user) { return (h && *(const int *)h == 1) ? cudns32_set_exchange(h, fn, user) : cudns64_set_exchange(h, fn, user); }
int cudns64_get_stream(cudns_handle h, void **stream);
int cudns32_get_stream(cudns_handle h, void **stream);
int cudns_get_stream(cudns_handle h, void **stream) { return (h && *(const int *)h == 1) ? cudns32_get_stream(h, stream) : cudns64_get_stream(h, stream); }
int cudns64_get_counters(cudns_handle h, uint64_t *kernel_launches, uint64_t *rk_stages);
int cudns32_get_counters(cudns_handle h, uint64_t *kernel_launches, uint64_t *rk_stages);
int cudns_get_counters(cudns_handle h, uint64_t *kernel_launches, uint64_t *rk_stages) { return (h && *(const int *)h == 1) ? cudns32_get_counters(h, kernel_launches, rk_stages) : cudns64_get_counters(h, kernel_launches, rk_stages); }
int cudns64_profile_stage(cudns_handle h, int reps, float *ms_theta, float *ms_rhs, float *ms_halo);
int cudns32_profile_stage(cudns_handle h, int reps, float *ms_theta, float *ms_rhs, float *ms_halo);
int cudns_profile_stage(cudns_handle h, int reps, float *ms_theta, float *ms_rhs, float *ms_halo) { return (h && *(const int *)h == 1) ? cudns32_profile_stage(h, reps, ms_theta, ms_rhs, ms_halo) : cudns64_profile_stage(h, reps, ms_theta, ms_rhs, ms_halo); }
int cudns64_set_stage_timing(cudns_handle h, int on);
int cudns32_set_stage_timing(cudns_handle h, int on);
int cudns_set_stage_timing(cudns_handle h, int on) { return (h && *(const int *)h == 1) ? cudns32_set_stage_timing(h, on) : cudns64_set_stage_timing(h, on); }
int cudns64_get_stage_timing(cudns_handle h, double *theta_ms, double *stage_ms, double *halo_ms, uint64_t *nstages);
int cudns32_get_stage_timing(cudns_handle h, double *theta_ms, double *stage_ms, double *halo_ms, uint64_t *nstages);
int cudns_get_stage_timing(cudns_handle h, double *theta_ms, double *stage_ms, double *halo_ms, uint64_t *nstages) { return (h && *(const int *)h == 1) ? cudns32_get_stage_timing(h, theta_ms, stage_ms, halo_ms, nstages) : cudns64_get_stage_timing(h, theta_ms, stage_ms, halo_ms, nstages); }
int cudns64_write_fields_async(cudns_handle h, const char *dir, int timestep);
int cudns32_write_fields_async(cudns_handle h, const char *dir, int timestep);
int cudns_write_fields_async(cudns_handle h, const char *dir, int timestep) { return (h && *(const int *)h == 1) ? cudns32_write_fields_async(h, dir, timestep) : cudns64_write_fields_async(h, dir, timestep); }
int cudns64_io_wait(cudns_handle h, uint64_t *files_written);
int cudns32_io_wait(cudns_handle h, uint64_t *files_written);
int cudns_io_wait(cudns_handle h, uint64_t *files_written) { return (h && *(const int *)h == 1) ? cudns32_io_wait(h, files_written) : cudns64_io_wait(h, files_written); }
int cudns64_read_fields(cudns_handle h, const char *dir, int timestep);
int cudns32_read_fields(cudns_handle h, const char *dir, int timestep);
int cudns_read_fields(cudns_handle h, const char *dir, int timestep) { return (h && *(const int *)h == 1) ? cudns32_read_fields(h, dir, timestep) : cudns64_read_fields(h, dir, timestep); }
int cudns64_calc_profiles(cudns_handle h, double *prof);
int cudns32_calc_profiles(cudns_handle h, double *prof);
int cudns_calc_profiles(cudns_handle h, double *prof) { return (h && *(const int *)h == 1) ? cudns32_calc_profiles(h, prof) : cudns64_calc_profiles(h, prof); }
int cudns64_calc_retau(cudns_handle h, double *retau);
int cudns32_calc_retau(cudns_handle h, double *retau);
int cudns_calc_retau(cudns_handle h, double *retau) { return (h && *(const int *)h == 1) ? cudns32_calc_retau(h, retau) : cudns64_calc_retau(h, retau); }
int cudns64_stats_begin(cudns_handle h, int nsnapshots);
int cudns32_stats_begin(cudns_handle h, int nsnapshots);
int cudns_stats_begin(cudns_handle h, int nsnapshots) { return (h && *(const int *)h == 1) ? cudns32_stats_begin(h, nsnapshots) : cudns64_stats_begin(h, nsnapshots); }
int cudns64_stats_add_mean(cudns_handle h);
int cudns32_stats_add_mean(cudns_handle h);
int cudns_stats_add_mean(cudns_handle h) { return (h && *(const int *)h == 1) ? cudns32_stats_add_mean(h) : cudns64_stats_add_mean(h); }
int cudns64_stats_finish_mean(cudns_handle h);
int cudns32_stats_finish_mean(cudns_handle h);
int cudns_stats_finish_mean(cudns_handle h) { return (h && *(const int *)h == 1) ? cudns32_stats_finish_mean(h) : cudns64_stats_finish_mean(h); }
int cudns64_stats_add_fluc(cudns_handle h);
int cudns32_stats_add_fluc(cudns_handle h);
int cudns_stats_add_fluc(cudns_handle h) { return (h && *(const int *)h == 1) ? cudns32_stats_add_fluc(h) : cudns64_stats_add_fluc(h); }
int cudns64_stats_get(cudns_handle h, double *mean, double *fluc, double *bulk, double *retau, double *utau);
int cudns32_stats_get(cudns_handle h, double *mean, double *fluc, double *bulk, double *retau, double *utau);
int cudns_stats_get(cudns_handle h, double *mean, double *fluc, double *bulk, double *retau, double *utau) { return (h && *(const int *)h == 1) ? cudns32_stats_get(h, mean, fluc, bulk, retau, utau) : cudns64_stats_get(h, mean, fluc, bulk, retau, utau); }
int cudns64_postprocess(cudns_handle h, const char *dir, int first, int last, const double *x, const char *outdir);
int cudns32_postprocess(cudns_handle h, const char *dir, int first, int last, const double *x, const char *outdir);
int cudns_postprocess(cudns_handle h, const char *dir, int first, int last, const double *x, const char *outdir) { return (h && *(const int *)h == 1) ? cudns32_postprocess(h, dir, first, last, x, outdir) : cudns64_postprocess(h, dir, first, last, x, outdir); }

}  // extern "C"
