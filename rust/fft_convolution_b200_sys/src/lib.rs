//! `impl Convolution` over the CUDA engine — the reference-side binding.
//!
//! The Rust host keeps exactly what the reference keeps on the host: the block scheduler, the
//! input-buffer fill and the segment-ring rotation of `FFTConvolver::process`
//! (reference src/fft_convolver.rs:222-231, :277-292); the arithmetic of every chunk
//! (forward FFT, delay-line MAC, inverse FFT + overlap-add) is the four `fcb_engine_*` stage
//! calls.  Contract violations surface as `panic!`, like the reference.
//!
//! This file is source only: it has not been compiled here (no Rust toolchain in the image).

use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct FcbEngine {
    _private: [u8; 0],
}

#[repr(C)]
pub struct FcbEngineDesc {
    pub channels: usize,
    pub block_size: usize,
    pub max_response_length: usize,
    pub shared_ir: c_int,
    pub device: c_int,
    pub stream: *mut c_void,
}

#[repr(C)]
pub struct FcbEpilogue {
    pub add0: *const f32,
    pub add1: *const f32,
    pub add_stride: usize,
    pub mix_other: *const f32,
    pub mix_stride: usize,
    pub gains: *const f32,
}

pub const FCB_OK: c_int = 0;

extern "C" {
    pub fn fcb_last_error() -> *const c_char;
    pub fn fcb_engine_create(desc: *const FcbEngineDesc, out: *mut *mut FcbEngine) -> c_int;
    pub fn fcb_engine_destroy(e: *mut FcbEngine);
    pub fn fcb_engine_clone(e: *const FcbEngine, out: *mut *mut FcbEngine) -> c_int;
    pub fn fcb_engine_block_size(e: *const FcbEngine) -> usize;
    pub fn fcb_engine_seg_count(e: *const FcbEngine) -> usize;
    pub fn fcb_engine_set_ir(e: *mut FcbEngine, chan0: usize, nchan: usize, irs: *const f32, len: usize,
                             stride: usize, is_update: c_int) -> c_int;
    pub fn fcb_engine_reset(e: *mut FcbEngine) -> c_int;
    pub fn fcb_engine_push_input(e: *mut FcbEngine, input: *const f32, stride: usize, fill: usize, n: usize) -> c_int;
    pub fn fcb_engine_fft_forward(e: *mut FcbEngine, current: usize, valid: usize) -> c_int;
    pub fn fcb_engine_mac(e: *mut FcbEngine, current: usize, active: usize) -> c_int;
    pub fn fcb_engine_ifft_ola(e: *mut FcbEngine, current: usize, fill: usize, n: usize, block_complete: c_int,
                               out_dev: *mut f32, out_stride: usize, epi: *const FcbEpilogue) -> c_int;
    pub fn fcb_engine_fetch(e: *mut FcbEngine, out_host: *mut f32, host_stride: usize, src_dev: *const f32,
                            dev_stride: usize, n: usize) -> c_int;
    pub fn fcb_engine_scratch(e: *mut FcbEngine) -> *mut f32;
    // several whole blocks in one time-batched pass (bit-identical to the block-by-block loop); host_io != 0: host pointers
    pub fn fcb_engine_multi_block_ok(e: *const FcbEngine, current: usize, active: usize) -> c_int;
    pub fn fcb_engine_multi_block_capacity(e: *mut FcbEngine) -> usize;
    pub fn fcb_engine_multi_block_reserve(e: *mut FcbEngine, nblocks: usize) -> c_int;
    pub fn fcb_engine_process_blocks(e: *mut FcbEngine, input: *const f32, in_stride: usize, output: *mut f32,
                                     out_stride: usize, current: usize, active: usize, nblocks: usize,
                                     epi: *const FcbEpilogue, host_io: c_int) -> c_int;
    // two engines fed the same input (two-stage head + tail0, crossfade A + B) in one launch
    pub fn fcb_engine_pair_ok(a: *const FcbEngine, b: *const FcbEngine, active: usize) -> c_int;
    pub fn fcb_engine_process_block_pair_dev(a: *mut FcbEngine, b: *mut FcbEngine, in_dev: *const f32, in_stride: usize,
                                             out_a: *mut f32, stride_a: usize, epi_a: *const FcbEpilogue,
                                             out_b: *mut f32, stride_b: usize, epi_b: *const FcbEpilogue,
                                             current: usize, active: usize) -> c_int;
}

fn check(rc: c_int) {
    if rc != FCB_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(fcb_last_error()) }.to_string_lossy().into_owned();
        panic!("{msg}");
    }
}

/// The reference's trait, src/lib.rs:5-14 (re-declared here so the crate stands alone; inside the
/// reference crate use `crate::Convolution` instead).
pub trait Convolution: Clone {
    fn init(response: &[f32], max_block_size: usize, max_response_length: usize) -> Self;
    fn update(&mut self, response: &[f32]);
    fn reset(&mut self);
    fn process(&mut self, input: &[f32], output: &mut [f32]);
}

/// Drop-in for `fft_convolver::FFTConvolver` (mono).  Host state = the reference's scalars
/// (src/fft_convolver.rs:88-91, :99, :101); everything else lives on the device.
pub struct CudaFFTConvolver {
    engine: *mut FcbEngine,
    ir_len: usize,
    block_size: usize,
    seg_count: usize,
    active_seg_count: usize,
    current: usize,
    input_buffer_fill: usize,
}

unsafe impl Send for CudaFFTConvolver {}

impl Convolution for CudaFFTConvolver {
    fn init(impulse_response: &[f32], block_size: usize, max_response_length: usize) -> Self {
        if max_response_length < impulse_response.len() {
            panic!("max_response_length must be at least the length of the initial impulse response");
        }
        let desc = FcbEngineDesc {
            channels: 1,
            block_size,
            max_response_length,
            shared_ir: 0,
            device: 0,
            stream: std::ptr::null_mut(),
        };
        let mut engine = std::ptr::null_mut();
        check(unsafe { fcb_engine_create(&desc, &mut engine) });
        // K5: segment FFTs of the zero-padded IR (src/fft_convolver.rs:131-142)
        check(unsafe {
            fcb_engine_set_ir(engine, 0, 1, impulse_response.as_ptr(), impulse_response.len(),
                              impulse_response.len(), 0)
        });
        let seg_count = unsafe { fcb_engine_seg_count(engine) };
        Self {
            engine,
            ir_len: max_response_length,
            block_size: unsafe { fcb_engine_block_size(engine) },
            seg_count,
            active_seg_count: seg_count,
            current: 0,
            input_buffer_fill: 0,
        }
    }

    // real-time safe: no allocation on either side of the boundary (staging is preallocated)
    fn update(&mut self, response: &[f32]) {
        if response.len() > self.ir_len {
            panic!("New impulse response is longer than initialized length");
        }
        if self.ir_len == 0 {
            return;
        }
        self.active_seg_count = (response.len() as f64 / self.block_size as f64).ceil() as usize;
        check(unsafe { fcb_engine_set_ir(self.engine, 0, 1, response.as_ptr(), response.len(), response.len(), 1) });
    }

    fn reset(&mut self) {
        check(unsafe { fcb_engine_reset(self.engine) });
        self.current = 0;
        self.input_buffer_fill = 0;
    }

    fn process(&mut self, input: &[f32], output: &mut [f32]) {
        if self.active_seg_count == 0 {
            output.fill(0.);
            return;
        }
        let scratch = unsafe { fcb_engine_scratch(self.engine) }; // device [1][B]
        let mut processed = 0;
        while processed < output.len() {
            let input_buffer_was_empty = self.input_buffer_fill == 0;
            // a call that spans several whole blocks: one time-batched pass over as many as the workspace holds
            let whole = (output.len() - processed) / self.block_size;
            if input_buffer_was_empty && whole >= 2
                && unsafe { fcb_engine_multi_block_ok(self.engine, self.current, self.active_seg_count) } != 0
            {
                let nb = std::cmp::min(whole, unsafe { fcb_engine_multi_block_capacity(self.engine) });
                let n = nb * self.block_size;
                let chunk = &input[processed..processed + n]; // panics like the reference if too short
                check(unsafe {
                    fcb_engine_process_blocks(self.engine, chunk.as_ptr(), n, output[processed..].as_mut_ptr(), n,
                                              self.current, self.active_seg_count, nb, std::ptr::null(), 1)
                });
                for _ in 0..nb {
                    self.current = if self.current > 0 { self.current - 1 } else { self.active_seg_count - 1 };
                }
                processed += n;
                continue;
            }
            let processing = std::cmp::min(output.len() - processed, self.block_size - self.input_buffer_fill);
            let pos = self.input_buffer_fill;
            let chunk = &input[processed..processed + processing]; // panics like the reference if too short
            unsafe {
                check(fcb_engine_push_input(self.engine, chunk.as_ptr(), processing, pos, processing));
                check(fcb_engine_fft_forward(self.engine, self.current, pos + processing)); // K1
                if input_buffer_was_empty {
                    check(fcb_engine_mac(self.engine, self.current, self.active_seg_count)); // K2
                }
                let complete = pos + processing == self.block_size;
                check(fcb_engine_ifft_ola(self.engine, self.current, pos, processing, complete as c_int,
                                          scratch, self.block_size, std::ptr::null())); // K3
                check(fcb_engine_fetch(self.engine, output[processed..].as_mut_ptr(), processing, scratch,
                                       self.block_size, processing));
            }
            self.input_buffer_fill += processing;
            if self.input_buffer_fill == self.block_size {
                self.input_buffer_fill = 0;
                self.current = if self.current > 0 { self.current - 1 } else { self.active_seg_count - 1 };
            }
            processed += processing;
        }
    }
}

impl Clone for CudaFFTConvolver {
    fn clone(&self) -> Self {
        let mut engine = std::ptr::null_mut();
        check(unsafe { fcb_engine_clone(self.engine, &mut engine) });
        Self { engine, ..*self }
    }
}

impl Drop for CudaFFTConvolver {
    fn drop(&mut self) {
        unsafe { fcb_engine_destroy(self.engine) };
    }
}
