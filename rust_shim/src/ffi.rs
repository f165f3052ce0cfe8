//! Raw bindings of include/bh_abi.h (ABI version 1), one declaration per exported symbol.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct bh_ctx { _p: [u8; 0] }
#[repr(C)] pub struct bh_ray_pipeline { _p: [u8; 0] }
#[repr(C)] pub struct bh_sky_pipeline { _p: [u8; 0] }
#[repr(C)] pub struct bh_post_pass { _p: [u8; 0] }
#[repr(C)] pub struct bh_frame_multi { _p: [u8; 0] }
#[repr(C)] pub struct bh_host_frame { _p: [u8; 0] }

pub const BH_ABI_VERSION: c_int = 2;
pub const BH_OK: c_int = 0;
pub const BH_ERR_NUMERIC: c_int = -34;
pub const BH_ERR_TIMEOUT: c_int = -110;
pub const BH_SHARED_FLAGS: usize = 64;
pub const BH_MODEL_UNIFORM_SIZE: usize = 48_234_572;

pub const BH_TEX_COLOR: c_int = 0;
pub const BH_TEX_DISK: c_int = 1;
pub const BH_TEX_SKY: c_int = 2;
pub const BH_SKY_RGBA16F: c_int = 0;
pub const BH_SKY_RGBA32F: c_int = 1;
pub const BH_POST_BLOOM_DOWN: c_int = 0;
pub const BH_POST_BLOOM_UP: c_int = 1;
pub const BH_POST_MIX: c_int = 2;
pub const BH_POST_HDR: c_int = 3;
pub const BH_POST_FXAA: c_int = 4;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct bh_pass_stats {
    pub ray_steps: u64, pub px_traced: u64, pub px_copied: u64, pub px_interp: u64,
    pub node_visits: u64, pub tri_tests: u64, pub tex_samples: u64, pub rk_reject: u64, pub stack_overflow: u64,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct bh_model_info {
    pub point_count: i32, pub normal_count: i32, pub triangle_count: i32,
    pub nodes_used: i32, pub max_depth: i32, pub leaf_count: i32, pub max_leaf_size: i32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct bh_frame_multi_desc {
    pub base_width: u32, pub base_height: u32, pub levels: u32, pub multiplier: u32, pub band_rows: u32, pub sky_format: i32,
}
#[repr(C)] #[derive(Clone, Copy)] pub struct bh_mix_details { pub mix_ratio: f32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct bh_fxaa_details { pub edge_threshold_min: f32, pub edge_threshold_max: f32, pub iterations: i32, pub subpixel_quality: f32 }

extern "C" {
    pub fn bh_abi_version() -> c_int;
    pub fn bh_last_error() -> *const c_char;

    pub fn bh_ctx_create(cuda_device: c_int, out: *mut *mut bh_ctx) -> c_int;
    pub fn bh_ctx_destroy(ctx: *mut bh_ctx);
    pub fn bh_ctx_set_numeric_mode(ctx: *mut bh_ctx, mode: c_int) -> c_int;
    pub fn bh_ctx_get_numeric_mode(ctx: *const bh_ctx) -> c_int;
    pub fn bh_ctx_set_texture(ctx: *mut bh_ctx, slot: c_int, rgba8: *const u8, w: u32, h: u32) -> c_int;
    pub fn bh_ctx_generate_disk_texture(ctx: *mut bh_ctx, w: u32, h: u32, host_rgba8: *mut u8, install: c_int) -> c_int;
    pub fn bh_ctx_upload_models(ctx: *mut bh_ctx, bytes: *const c_void, nbytes: usize) -> c_int;
    pub fn bh_ctx_upload_models_async(ctx: *mut bh_ctx, pinned_bytes: *const c_void, nbytes: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn bh_ctx_set_model_header(ctx: *mut bh_ctx, index: u32, position: *const f32, visible: i32) -> c_int;

    pub fn bh_ray_pipeline_create(ctx: *mut bh_ctx, width: u32, height: u32, prev: *const bh_ray_pipeline,
                                  out: *mut *mut bh_ray_pipeline) -> c_int;
    pub fn bh_ray_pipeline_destroy(p: *mut bh_ray_pipeline);
    pub fn bh_ray_pipeline_set_tiling(p: *mut bh_ray_pipeline, band_rows: u32, rank: u32, n_ranks: u32) -> c_int;
    pub fn bh_ray_pipeline_local_rows(p: *const bh_ray_pipeline) -> u32;
    pub fn bh_ray_pipeline_enable_aux(p: *mut bh_ray_pipeline, aux_mask: u32) -> c_int;
    pub fn bh_ray_pipeline_bind_output(p: *mut bh_ray_pipeline, device_rgba32f: *mut c_void) -> c_int;
    pub fn bh_ray_pipeline_bind_frame(p: *mut bh_ray_pipeline, device_frame_rgba32f: *mut c_void) -> c_int;
    pub fn bh_shared_frame_create(ctx: *mut bh_ctx, nbytes: usize, device_ptr: *mut *mut c_void, handle_out: *mut u8) -> c_int;
    pub fn bh_shared_frame_open(ctx: *mut bh_ctx, handle: *const u8, device_ptr: *mut *mut c_void) -> c_int;
    pub fn bh_shared_frame_release(ctx: *mut bh_ctx, device_ptr: *mut c_void, owner: c_int) -> c_int;
    pub fn bh_ray_pipeline_pass(p: *mut bh_ray_pipeline, camera: *const c_void, black_hole: *const c_void,
                                details: *const c_void, cuda_stream: *mut c_void) -> c_int;
    pub fn bh_ray_pipeline_pass_to_host(p: *mut bh_ray_pipeline, camera: *const c_void, black_hole: *const c_void,
                                        details: *const c_void, pinned_host_rgba32f: *mut f32, n_chunks: u32,
                                        cuda_stream: *mut c_void) -> c_int;
    pub fn bh_ray_pipeline_sync(p: *mut bh_ray_pipeline) -> c_int;
    pub fn bh_ray_pipeline_output(p: *const bh_ray_pipeline) -> *const f32;
    pub fn bh_ray_pipeline_width(p: *const bh_ray_pipeline) -> u32;
    pub fn bh_ray_pipeline_height(p: *const bh_ray_pipeline) -> u32;
    pub fn bh_ray_pipeline_read(p: *mut bh_ray_pipeline, host_rgba32f: *mut f32, host_hit: *mut i32,
                                host_steps: *mut u32, host_class: *mut u8) -> c_int;
    pub fn bh_ray_pipeline_stats(p: *mut bh_ray_pipeline, out: *mut bh_pass_stats) -> c_int;

    pub fn bh_sky_pipeline_create(ctx: *mut bh_ctx, prev: *const bh_ray_pipeline, format: c_int, out: *mut *mut bh_sky_pipeline) -> c_int;
    pub fn bh_sky_pipeline_create_for_frame(ctx: *mut bh_ctx, device_frame_rgba32f: *const c_void, width: u32, height: u32,
                                            format: c_int, out: *mut *mut bh_sky_pipeline) -> c_int;
    pub fn bh_sky_pipeline_destroy(p: *mut bh_sky_pipeline);
    pub fn bh_sky_pipeline_bind_output(p: *mut bh_sky_pipeline, device_rgba: *mut c_void) -> c_int;
    pub fn bh_sky_pipeline_pass(p: *mut bh_sky_pipeline, cuda_stream: *mut c_void) -> c_int;
    pub fn bh_sky_pipeline_output(p: *const bh_sky_pipeline) -> *const c_void;
    pub fn bh_sky_pipeline_read(p: *mut bh_sky_pipeline, host_rgba: *mut c_void) -> c_int;

    pub fn bh_frame_multi_create(ctxs: *const *mut bh_ctx, n_devices: u32, desc: *const bh_frame_multi_desc,
                                 out: *mut *mut bh_frame_multi) -> c_int;
    pub fn bh_frame_multi_destroy(fm: *mut bh_frame_multi);
    pub fn bh_frame_multi_width(fm: *const bh_frame_multi) -> u32;
    pub fn bh_frame_multi_height(fm: *const bh_frame_multi) -> u32;
    pub fn bh_frame_multi_pass(fm: *mut bh_frame_multi, camera: *const c_void, black_hole: *const c_void, details: *const c_void) -> c_int;
    pub fn bh_frame_multi_pass_to_host(fm: *mut bh_frame_multi, camera: *const c_void, black_hole: *const c_void,
                                       details: *const c_void, pinned_host_rgba32f: *mut f32) -> c_int;
    pub fn bh_frame_multi_sync(fm: *mut bh_frame_multi) -> c_int;
    pub fn bh_frame_multi_output(fm: *const bh_frame_multi) -> *const f32;
    pub fn bh_frame_multi_sky_output(fm: *const bh_frame_multi) -> *const c_void;
    pub fn bh_frame_multi_read(fm: *mut bh_frame_multi, host_rgba32f: *mut f32, host_sky: *mut c_void) -> c_int;
    pub fn bh_frame_multi_stats(fm: *mut bh_frame_multi, out: *mut bh_pass_stats, elapsed_ms: *mut f32) -> c_int;
    pub fn bh_shared_frame_flags(device_ptr: *mut c_void, nbytes: usize) -> *mut u32;
    pub fn bh_stream_signal(ctx: *mut bh_ctx, device_flag: *mut u32, value: u32, cuda_stream: *mut c_void) -> c_int;
    pub fn bh_stream_wait(ctx: *mut bh_ctx, device_flags: *const u32, n_flags: u32, value: u32, timeout_ms: u32,
                          cuda_stream: *mut c_void) -> c_int;
    pub fn bh_ctx_check_async(ctx: *mut bh_ctx) -> c_int;
    pub fn bh_host_frame_create(ctx: *mut bh_ctx, shm_name: *const c_char, nbytes: usize, create: c_int, out: *mut *mut bh_host_frame) -> c_int;
    pub fn bh_host_frame_destroy(hf: *mut bh_host_frame, unlink_name: c_int);
    pub fn bh_host_frame_ptr(hf: *const bh_host_frame) -> *mut f32;
    pub fn bh_host_frame_signal(hf: *mut bh_host_frame, slot: u32, value: u32) -> c_int;
    pub fn bh_host_frame_wait(hf: *mut bh_host_frame, first_slot: u32, n_slots: u32, value: u32, timeout_ms: u32) -> c_int;
    pub fn bh_ray_pipeline_pass_to_host_frame(p: *mut bh_ray_pipeline, camera: *const c_void, black_hole: *const c_void,
                                              details: *const c_void, mapped_host_frame_rgba32f: *mut f32,
                                              cuda_stream: *mut c_void) -> c_int;

    pub fn bh_post_pass_create(ctx: *mut bh_ctx, kind: c_int, out_w: u32, out_h: u32, in1_device: *const c_void, in1_w: u32,
                               in1_h: u32, in2_device: *const c_void, out: *mut *mut bh_post_pass) -> c_int;
    pub fn bh_post_pass_destroy(p: *mut bh_post_pass);
    pub fn bh_post_pass_run(p: *mut bh_post_pass, details: *const c_void, cuda_stream: *mut c_void) -> c_int;
    pub fn bh_post_pass_output(p: *const bh_post_pass) -> *const c_void;
    pub fn bh_post_pass_read(p: *mut bh_post_pass, host: *mut c_void) -> c_int;

    pub fn bh_model_validate(model_uniform: *const c_void) -> c_int;
    pub fn bh_model_load_obj(path: *const c_char, model_uniform: *mut c_void, info: *mut bh_model_info) -> c_int;
    pub fn bh_model_from_arrays(points: *const f32, n_points: i32, normals: *const f32, n_normals: i32, tris: *const i32,
                                n_tris: i32, position: *const f32, visible: i32, model_uniform: *mut c_void,
                                info: *mut bh_model_info) -> c_int;
    pub fn bh_model_build_bvh(model_uniform: *mut c_void, triangle_count: i32, info: *mut bh_model_info) -> c_int;
    pub fn bh_save_png(path: *const c_char, rgba8: *const u8, w: u32, h: u32, force_opaque: c_int) -> c_int;
    pub fn bh_ctx_math_probe(ctx: *mut bh_ctx, func: c_int, host_a: *const f32, host_b: *const f32, host_out: *mut f32, n: usize) -> c_int;
}
