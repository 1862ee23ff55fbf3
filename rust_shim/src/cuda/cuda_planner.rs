//! `src/cuda/cuda_planner.rs` -- `FftPlannerCuda<T>`, shaped like `FftPlannerSse<T>`
//! (`src/sse/sse_planner.rs:144-226`) and `FftPlannerAvx<T>` (`src/avx/avx_planner.rs:113-185`).
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image).

use std::any::TypeId;
use std::os::raw::c_int;
use std::sync::Arc;

use super::{b200fft_device_count, B200FftRecipeNode, CudaFft};
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
        // Rust owns the decomposition: the scalar planner's recipe (src/plan.rs:412-425) is flattened and handed to the
        // library; shapes it has no kernel sequence for (e.g. Rader over a non-smooth inner length) fall back to the
        // library's OWN choice for the same length.  There is no CPU fallback on this path (north_star): a length the
        // library cannot plan at all (non-power-of-two above 2^23, power of two above 2^24) panics with its message.
        let recipe = flatten_recipe(&crate::plan::FftPlannerScalar::<T>::new().design_fft_for_len(len));
        let planned = CudaFft::<T>::from_recipe(&recipe, direction, self.device).or_else(|| CudaFft::<T>::new(len, direction, self.device));
        let fft: Arc<dyn Fft<T>> = match planned {
            Some(f) => Arc::new(f),
            None => panic!("FftPlannerCuda: {}", super::last_error_text()),
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

/// `crate::plan::Recipe` (src/plan.rs:134-226) -> the node array of include/b200fft.h.  Leaf butterflies and Radix4 / RadixN
/// collapse into POW2 / SMOOTH nodes (the library picks its own radices inside one pass); MixedRadix / GoodThomas keep their
/// split, Rader / Bluestein keep their inner FFT as a child node.
fn flatten_recipe(recipe: &crate::plan::Recipe) -> Vec<B200FftRecipeNode> {
    use crate::plan::Recipe::*;
    fn push(out: &mut Vec<B200FftRecipeNode>, r: &crate::plan::Recipe) -> u32 {
        let idx = out.len() as u32;
        let len = r.len() as u64;
        out.push(B200FftRecipeNode { kind: 0, child: 0, len, a: 0, b: 0 });
        match r {
            MixedRadix { left_fft, right_fft } | MixedRadixSmall { left_fft, right_fft } => {
                out[idx as usize].kind = 3;
                out[idx as usize].a = left_fft.len() as u64;
                out[idx as usize].b = right_fft.len() as u64;
            }
            GoodThomasAlgorithm { left_fft, right_fft } | GoodThomasAlgorithmSmall { left_fft, right_fft } => {
                out[idx as usize].kind = 4;
                out[idx as usize].a = left_fft.len() as u64;
                out[idx as usize].b = right_fft.len() as u64;
            }
            RadersAlgorithm { inner_fft } => {
                out[idx as usize].kind = 5;
                let c = push(out, inner_fft);
                out[idx as usize].child = c;
            }
            BluesteinsAlgorithm { inner_fft, .. } => {
                out[idx as usize].kind = 6;
                let c = push(out, inner_fft);
                out[idx as usize].child = c;
            }
            _ => {
                out[idx as usize].kind = if (len & (len.wrapping_sub(1))) == 0 { 1 } else { 2 };
            }
        }
        idx
    }
    let mut out = Vec::new();
    push(&mut out, recipe);
    out
}
