// libcudns: names of the precision-specific copies of the device-side C entry points (everything include/cudns.h declares that needs a
// solver handle).  api.cu is compiled twice (real = double / float, see cudns_internal.h); this header, included BEFORE
// include/cudns.h, renames its definitions -- and the solver object -- to cudns64_* / cudns32_*.  abi_dispatch.cpp defines the public
// symbols and forwards to the copy the handle belongs to.
#pragma once
#ifdef CUDNS_F32
#define CUDNS_PREC_NAME(x) cudns32_##x
#else
#define CUDNS_PREC_NAME(x) cudns64_##x
#endif
#define cudns_solver CUDNS_PREC_NAME(solver)
#define cudns_create CUDNS_PREC_NAME(create)
#define cudns_destroy CUDNS_PREC_NAME(destroy)
#define cudns_memory_report CUDNS_PREC_NAME(memory_report)
#define cudns_set_state CUDNS_PREC_NAME(set_state)
#define cudns_get_state CUDNS_PREC_NAME(get_state)
#define cudns_set_state_device CUDNS_PREC_NAME(set_state_device)
#define cudns_get_state_device CUDNS_PREC_NAME(get_state_device)
#define cudns_set_sponge CUDNS_PREC_NAME(set_sponge)
#define cudns_advance CUDNS_PREC_NAME(advance)
#define cudns_calc_rhs CUDNS_PREC_NAME(calc_rhs)
#define cudns_calc_dt CUDNS_PREC_NAME(calc_dt)
#define cudns_calc_bulk CUDNS_PREC_NAME(calc_bulk)
#define cudns_calc_enstrophy CUDNS_PREC_NAME(calc_enstrophy)
#define cudns_get_scalars CUDNS_PREC_NAME(get_scalars)
#define cudns_set_dt CUDNS_PREC_NAME(set_dt)
#define cudns_halo_local_info CUDNS_PREC_NAME(halo_local_info)
#define cudns_halo_connect CUDNS_PREC_NAME(halo_connect)
#define cudns_halo_buffers CUDNS_PREC_NAME(halo_buffers)
#define cudns_set_allreduce CUDNS_PREC_NAME(set_allreduce)
#define cudns_set_exchange CUDNS_PREC_NAME(set_exchange)
#define cudns_get_stream CUDNS_PREC_NAME(get_stream)
#define cudns_get_counters CUDNS_PREC_NAME(get_counters)
#define cudns_profile_stage CUDNS_PREC_NAME(profile_stage)
#define cudns_set_stage_timing CUDNS_PREC_NAME(set_stage_timing)
#define cudns_get_stage_timing CUDNS_PREC_NAME(get_stage_timing)
#define cudns_write_fields_async CUDNS_PREC_NAME(write_fields_async)
#define cudns_io_wait CUDNS_PREC_NAME(io_wait)
#define cudns_read_fields CUDNS_PREC_NAME(read_fields)
#define cudns_calc_profiles CUDNS_PREC_NAME(calc_profiles)
#define cudns_calc_retau CUDNS_PREC_NAME(calc_retau)
#define cudns_stats_begin CUDNS_PREC_NAME(stats_begin)
#define cudns_stats_add_mean CUDNS_PREC_NAME(stats_add_mean)
#define cudns_stats_finish_mean CUDNS_PREC_NAME(stats_finish_mean)
#define cudns_stats_add_fluc CUDNS_PREC_NAME(stats_add_fluc)
#define cudns_stats_get CUDNS_PREC_NAME(stats_get)
#define cudns_postprocess CUDNS_PREC_NAME(postprocess)
