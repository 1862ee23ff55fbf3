//! `src/cuda/cuda_planner.rs` -- `FftPlannerCuda<T>`, shaped like `FftPlannerSse<T>`
//! (`src/sse/sse_planner.rs:144-226`) and `FftPlannerAvx<T>` (`src/avx/avx_planner.rs:113-185`).
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image).

use std::any::TypeId;
use std::os::raw::c_int;
use std::sync::Arc;

use super::{b200fft_device_count, CudaFft};
use crate::common::FftNum;
use crate::fft_cache::FftCache;
use crate::{Fft, FftDirection};

pub struct FftPlannerCuda<T: FftNum> {
    cache: FftCache<T>, // one instance per (len, direction), src/fft_cache.rs:5-38
    device: i32,
}

impl<T: FftNum> FftPlannerCuda<T> {
    /// `Err(())` when no B200 is visible or `T` is not f32/f64 -- `FftPlanner::new()` then falls
    /// through to the next backend exactly as it does for AVX -> SSE -> NEON (src/plan.rs:72-94).
    pub fn new() -> Result<Self, ()> {
        let is_float = TypeId::of::<T>() == TypeId::of::<f32>() || TypeId::of::<T>() == TypeId::of::<f64>();
        let mut n: c_int = 0;
        let rc = unsafe { b200fft_device_count(&mut n) };
        if rc == 0 && n > 0 && is_float {
            Ok(Self { cache: FftCache::new(), device: 0 })
        } else {
            Err(())
        }
    }

    pub fn plan_fft(&mut self, len: usize, direction: FftDirection) -> Arc<dyn Fft<T>> {
        if let Some(instance) = self.cache.get(len, direction) {
            return instance;
        }
        let fft: Arc<dyn Fft<T>> = match CudaFft::<T>::new(len, direction, self.device) {
            Some(f) => Arc::new(f),
            // lengths this build of libb200fft does not plan: hand them to the scalar planner so the
            // planner as a whole still returns an FFT for every length (RustFFT's planners never fail)
            None => crate::FftPlannerScalar::<T>::new().plan_fft(len, direction),
        };
        self.cache.insert(&fft);
        fft
    }
    pub fn plan_fft_forward(&mut self, len: usize) -> Arc<dyn Fft<T>> {
        self.plan_fft(len, FftDirection::Forward)
    }
    pub fn plan_fft_inverse(&mut self, len: usize) -> Arc<dyn Fft<T>> {
        self.plan_fft(len, FftDirection::Inverse)
    }
}
