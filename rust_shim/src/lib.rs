//! Pass objects with the method names of bhusie's `src/renderer/pipelines/` (new / output_view / pass), backed by
//! libbhray.so.  What `Renderer::new` (src/renderer/mod.rs:170-216) and `Renderer::render` (mod.rs:406-421) would hold
//! instead of `RayPipeline` / `SkyPipeline`; see INTEGRATION.md §3 for the call-site changes.
//!
//! The uniforms are passed per frame as the `#[repr(C)]` structs the renderer already owns (`CameraUniform`,
//! `BlackHoleUniform`, `RayDetails`): their bytes are exactly what it hands to `queue.write_buffer` (mod.rs:386-388).
//! Like the reference's constructors (`unwrap` / `expect`), failures panic with the library's error text.
pub mod ffi;

use std::ffi::CStr;
use std::os::raw::c_void;
use std::ptr;

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::bh_last_error()) }.to_string_lossy().into_owned()
}
fn check(rc: i32, what: &str) {
    assert!(rc == ffi::BH_OK, "{what}: libbhray error {rc}: {}", last_error());
}

/// One per CUDA device: owns the three textures (`texture.rs`, decoded once instead of once per pipeline as
/// ray_pipeline.rs:63-70 does) and the `ModelUniform` array (`array_buffer.rs`).
pub struct CudaContext { raw: *mut ffi::bh_ctx }

impl CudaContext {
    pub fn new(cuda_device: i32) -> Self {
        assert_eq!(unsafe { ffi::bh_abi_version() }, ffi::BH_ABI_VERSION, "libbhray ABI version");
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::bh_ctx_create(cuda_device, &mut raw) }, "bh_ctx_create");
        Self { raw }
    }
    /// `slot`: ffi::BH_TEX_COLOR / BH_TEX_DISK / BH_TEX_SKY; `rgba8`: `image::load_from_memory(..).to_rgba8()` bytes.
    pub fn set_texture(&self, slot: i32, rgba8: &[u8], width: u32, height: u32) {
        assert_eq!(rgba8.len(), width as usize * height as usize * 4);
        check(unsafe { ffi::bh_ctx_set_texture(self.raw, slot, rgba8.as_ptr(), width, height) }, "bh_ctx_set_texture");
    }
    /// `bytemuck::bytes_of(&model_uniform)` for each model, concatenated (once at load, not per frame).
    pub fn upload_models(&self, model_uniform_bytes: &[u8]) {
        check(unsafe { ffi::bh_ctx_upload_models(self.raw, model_uniform_bytes.as_ptr() as *const c_void, model_uniform_bytes.len()) },
              "bh_ctx_upload_models");
    }
    /// The per-frame part of `scene.models.update_buffer` (array_buffer.rs:71-79): position and visibility only.
    pub fn set_model_header(&self, index: u32, position: [f32; 3], visible: bool) {
        check(unsafe { ffi::bh_ctx_set_model_header(self.raw, index, position.as_ptr(), visible as i32) }, "bh_ctx_set_model_header");
    }
}
impl Drop for CudaContext { fn drop(&mut self) { unsafe { ffi::bh_ctx_destroy(self.raw) } } }

/// `RayPipeline` (ray_pipeline.rs:28-33).  `prev == None` is the 1x1 base texture of mod.rs:151-168.
pub struct CudaRayPipeline { raw: *mut ffi::bh_ray_pipeline, pub resolution: (u32, u32) }

impl CudaRayPipeline {
    pub fn new(ctx: &CudaContext, resolution: (u32, u32), prev: Option<&CudaRayPipeline>) -> Self {
        let mut raw = ptr::null_mut();
        let prev_raw = prev.map_or(ptr::null(), |p| p.raw as *const ffi::bh_ray_pipeline);
        check(unsafe { ffi::bh_ray_pipeline_create(ctx.raw, resolution.0, resolution.1, prev_raw, &mut raw) }, "bh_ray_pipeline_create");
        Self { raw, resolution }
    }
    /// `RayPipeline::output_view`: a device pointer (row-major RGBA32F) instead of a `TextureView`.
    pub fn output_view(&self) -> *const f32 { unsafe { ffi::bh_ray_pipeline_output(self.raw) } }
    /// `RayPipeline::pass`: enqueues on the default stream and returns, like recording into a `ComputePass`.
    pub fn pass<C: bytemuck::Pod, B: bytemuck::Pod, D: bytemuck::Pod>(&mut self, camera: &C, black_hole: &B, details: &D) {
        assert_eq!((std::mem::size_of::<C>(), std::mem::size_of::<B>(), std::mem::size_of::<D>()), (32, 132, 32));
        check(unsafe { ffi::bh_ray_pipeline_pass(self.raw, bytemuck::bytes_of(camera).as_ptr() as *const c_void,
                                                 bytemuck::bytes_of(black_hole).as_ptr() as *const c_void,
                                                 bytemuck::bytes_of(details).as_ptr() as *const c_void, ptr::null_mut()) },
              "bh_ray_pipeline_pass");
    }
    pub fn stats(&mut self) -> ffi::bh_pass_stats {
        let mut s = ffi::bh_pass_stats::default();
        let rc = unsafe { ffi::bh_ray_pipeline_stats(self.raw, &mut s) };
        assert!(rc == ffi::BH_OK || rc == ffi::BH_ERR_NUMERIC, "bh_ray_pipeline_stats: {}", last_error());
        s
    }
}
impl Drop for CudaRayPipeline { fn drop(&mut self) { unsafe { ffi::bh_ray_pipeline_destroy(self.raw) } } }

/// `SkyPipeline` (sky_pipeline.rs:10-16): Rgba16Float output like the reference.
pub struct CudaSkyPipeline { raw: *mut ffi::bh_sky_pipeline, pub resolution: (u32, u32) }

impl CudaSkyPipeline {
    pub fn new(ctx: &CudaContext, prev: &CudaRayPipeline) -> Self {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::bh_sky_pipeline_create(ctx.raw, prev.raw, ffi::BH_SKY_RGBA16F, &mut raw) }, "bh_sky_pipeline_create");
        Self { raw, resolution: prev.resolution }
    }
    pub fn output_view(&self) -> *const c_void { unsafe { ffi::bh_sky_pipeline_output(self.raw) } }
    pub fn pass(&mut self) { check(unsafe { ffi::bh_sky_pipeline_pass(self.raw, ptr::null_mut()) }, "bh_sky_pipeline_pass"); }
    /// Host staging for the wgpu post chain: width*height*4 half floats, synchronous.
    pub fn read(&mut self, host_rgba16f: &mut [u16]) {
        assert_eq!(host_rgba16f.len(), self.resolution.0 as usize * self.resolution.1 as usize * 4);
        check(unsafe { ffi::bh_sky_pipeline_read(self.raw, host_rgba16f.as_mut_ptr() as *mut c_void) }, "bh_sky_pipeline_read");
    }
}
impl Drop for CudaSkyPipeline { fn drop(&mut self) { unsafe { ffi::bh_sky_pipeline_destroy(self.raw) } } }

/// The pyramid of mod.rs:170-207: base (72, 41), each level 3x - 2, four levels; returns the levels and the sky resolve.
pub fn build_pyramid(ctx: &CudaContext, base: (f32, f32), multiplier: f32, iterations: usize) -> (Vec<CudaRayPipeline>, CudaSkyPipeline) {
    let mut levels: Vec<CudaRayPipeline> = Vec::with_capacity(iterations);
    let mut curr = base;
    for _ in 0..iterations {
        let level = CudaRayPipeline::new(ctx, (curr.0 as u32, curr.1 as u32), levels.last());
        levels.push(level);
        curr = (curr.0 * multiplier - (multiplier - 1.0), curr.1 * multiplier - (multiplier - 1.0));
    }
    let sky = CudaSkyPipeline::new(ctx, levels.last().expect("at least one level"));
    (levels, sky)
}

/// The whole compute pass of `Renderer::render` (mod.rs:406-421) on N GPUs from this one thread: coarse pyramid levels
/// replicated per device, the last level cut into cyclic row bands whose pixels every device stores straight into device 0's
/// frame over NVLink, sky resolve on device 0 (`bh_frame_multi`, include/bh_abi.h).  Each context must already hold the
/// textures and models.
pub struct CudaFrameMulti { raw: *mut ffi::bh_frame_multi, pub resolution: (u32, u32) }

impl CudaFrameMulti {
    /// `base`, `multiplier`, `iterations` as in mod.rs:177-179 ((72, 41), 3, 4); `iterations == 1` renders `base` itself.
    pub fn new(ctxs: &[&CudaContext], base: (u32, u32), multiplier: u32, iterations: u32, band_rows: u32) -> Self {
        let raws: Vec<*mut ffi::bh_ctx> = ctxs.iter().map(|c| c.raw).collect();
        let desc = ffi::bh_frame_multi_desc { base_width: base.0, base_height: base.1, levels: iterations, multiplier, band_rows,
                                              sky_format: ffi::BH_SKY_RGBA16F };
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::bh_frame_multi_create(raws.as_ptr(), raws.len() as u32, &desc, &mut raw) }, "bh_frame_multi_create");
        let resolution = unsafe { (ffi::bh_frame_multi_width(raw), ffi::bh_frame_multi_height(raw)) };
        Self { raw, resolution }
    }
    /// all ray levels + the sky resolve, enqueued on the object's own streams; returns at once like `ComputePass` recording
    pub fn pass<C: bytemuck::Pod, B: bytemuck::Pod, D: bytemuck::Pod>(&mut self, camera: &C, black_hole: &B, details: &D) {
        assert_eq!((std::mem::size_of::<C>(), std::mem::size_of::<B>(), std::mem::size_of::<D>()), (32, 132, 32));
        check(unsafe { ffi::bh_frame_multi_pass(self.raw, bytemuck::bytes_of(camera).as_ptr() as *const c_void,
                                                bytemuck::bytes_of(black_hole).as_ptr() as *const c_void,
                                                bytemuck::bytes_of(details).as_ptr() as *const c_void) }, "bh_frame_multi_pass");
    }
    /// `SkyPipeline::output_view` of the assembled frame: device pointer (device 0), Rgba16Float
    pub fn output_view(&self) -> *const c_void { unsafe { ffi::bh_frame_multi_sky_output(self.raw) } }
    pub fn read(&mut self, host_rgba16f: &mut [u16]) {
        assert_eq!(host_rgba16f.len(), self.resolution.0 as usize * self.resolution.1 as usize * 4);
        check(unsafe { ffi::bh_frame_multi_read(self.raw, ptr::null_mut(), host_rgba16f.as_mut_ptr() as *mut c_void) }, "bh_frame_multi_read");
    }
}
impl Drop for CudaFrameMulti { fn drop(&mut self) { unsafe { ffi::bh_frame_multi_destroy(self.raw) } } }
