//! Rust side of the FFI seam (UNCOMPILED — no rustc in the build image).
//!
//! `ffi` mirrors include/qsv.h one to one.  `DeviceState` is the owner type that replaces the
//! `register: SuperPosition` field of quantr's `SimulatedCircuit` (src/simulated_circuit.rs:20-27);
//! `encode` walks `circuit_gates` exactly like `simulate_with_register` (src/circuit/simulation.rs:37-56)
//! and expands `Gate::Custom` closures on the 2^k basis states (src/circuit/simulation.rs:137-156).
//! INTEGRATION.md shows where these are called from inside quantr.

use num_complex::Complex64;
use std::os::raw::{c_char, c_int, c_void};

pub mod ffi {
    use super::*;

    #[repr(C)]
    pub struct QsvOp {
        pub kind: u32,
        pub target: u32,
        pub n_controls: u32,
        pub reserved: u32,
        pub controls: *const u32,
        pub param: f64,
        pub iparam: i32,
        pub reserved2: i32,
        pub matrix: *const f64,
        pub none_mask: *const u8,
    }

    #[repr(C)]
    #[derive(Default)]
    pub struct QsvStats {
        pub n_gates: u64,
        pub n_passes: u64,
        pub n_rounds: u64,
        pub n_kernel_launches: u64,
        pub bytes_per_pass: u64,
        pub n_exchanges: u64,
        pub exchange_bytes: u64,
        pub device_ms: f64,
        pub exchange_ms: f64,
    }

    #[repr(C)]
    pub struct QsvState {
        _private: [u8; 0],
    }

    #[repr(C)]
    pub struct QsvPlan {
        _private: [u8; 0],
    }

    // Every export of include/qsv.h, in the header's order (tests/test_abi.py checks the library exports the same list).
    extern "C" {
        pub fn qsv_create(out: *mut *mut QsvState, n_qubits: u32, device: c_int) -> c_int;
        pub fn qsv_create_sharded(out: *mut *mut QsvState, n_qubits: u32, device: c_int, rank: c_int, world: c_int,
                                  nccl_unique_id: *const c_void, nccl_unique_id_bytes: usize) -> c_int;
        pub fn qsv_create_multi(out: *mut *mut QsvState, n_qubits: u32, devices: *const c_int, n_devices: c_int) -> c_int;
        pub fn qsv_nccl_unique_id(out: *mut c_void, out_bytes: usize) -> c_int;
        pub fn qsv_peer_export(s: *mut QsvState, out_handle: *mut c_void, out_bytes: usize) -> c_int;
        pub fn qsv_peer_import(s: *mut QsvState, handles: *const c_void, n_handles: usize) -> c_int;
        pub fn qsv_destroy(s: *mut QsvState) -> c_int;
        pub fn qsv_last_error(s: *const QsvState) -> *const c_char;
        pub fn qsv_set_option(s: *mut QsvState, key: *const c_char, value: i64) -> c_int;
        pub fn qsv_get_info(s: *const QsvState, key: *const c_char, value: *mut i64) -> c_int;
        pub fn qsv_init_basis(s: *mut QsvState, index: u64) -> c_int;
        pub fn qsv_upload(s: *mut QsvState, host_amps: *const f64, first: u64, count: u64) -> c_int;
        pub fn qsv_download(s: *mut QsvState, host_amps: *mut f64, first: u64, count: u64) -> c_int;
        pub fn qsv_gather(s: *mut QsvState, indices: *const u64, count: u64, host_amps: *mut f64) -> c_int;
        pub fn qsv_apply(s: *mut QsvState, ops: *const QsvOp, n_ops: usize, stats: *mut QsvStats) -> c_int;
        pub fn qsv_plan_create(out: *mut *mut QsvPlan, n_qubits: u32, n_local_qubits: u32, ops: *const QsvOp, n_ops: usize,
                               tile_bits: u32, low_bits: u32, fuse: c_int) -> c_int;
        pub fn qsv_plan_create_ex(out: *mut *mut QsvPlan, n_qubits: u32, n_local_qubits: u32, ops: *const QsvOp, n_ops: usize,
                                  tile_bits: u32, low_bits: u32, fuse: c_int, layout: *const u8, free_layout: c_int) -> c_int;
        pub fn qsv_plan_destroy(p: *mut QsvPlan) -> c_int;
        pub fn qsv_plan_initial_amplitudes(p: *const QsvPlan, basis_index: u64, out: *mut f64, cap: usize) -> c_int;
        pub fn qsv_plan_num_steps(p: *const QsvPlan, n_steps: *mut usize) -> c_int;
        pub fn qsv_plan_get_step(p: *const QsvPlan, i: usize, kind: *mut c_int, pass_index: *mut u32, partner_bits: *mut u8, cap: usize) -> c_int;
        pub fn qsv_plan_get_layout(p: *const QsvPlan, which: c_int, out_layout: *mut u8, cap: usize) -> c_int;
        pub fn qsv_plan_stats(p: *const QsvPlan, stats: *mut QsvStats) -> c_int;
        pub fn qsv_plan_serialize(p: *const QsvPlan, out: *mut c_void, cap: usize, size: *mut usize) -> c_int;
        pub fn qsv_plan_last_error() -> *const c_char;
        pub fn qsv_run_plan(s: *mut QsvState, p: *mut QsvPlan, stats: *mut QsvStats) -> c_int;
        pub fn qsv_sample(s: *mut QsvState, uniforms: *const f64, shots: u64, out_indices: *mut u64) -> c_int;
        pub fn qsv_norm_sqr(s: *mut QsvState, out: *mut f64) -> c_int;
        pub fn qsv_get_layout(s: *const QsvState, out_layout: *mut u8, cap: usize) -> c_int;
        pub fn qsv_last_step_ms(s: *const QsvState, out_ms: *mut f64, cap: usize, n_steps: *mut usize) -> c_int;
        pub fn qsv_save(s: *mut QsvState, path: *const c_char) -> c_int;
        pub fn qsv_load(s: *mut QsvState, path: *const c_char) -> c_int;
        pub fn qsv_synchronize(s: *mut QsvState) -> c_int;
        pub fn qsv_device_pointer(s: *mut QsvState, dev_ptr: *mut *mut c_void, cuda_stream: *mut *mut c_void) -> c_int;
    }

    pub const GATE_CUSTOM: u32 = 24; // QSV_GATE_* follow the declaration order of quantr's `enum Gate`
}

/// Owns a `qsv_state*`.  `!Sync`: one call at a time per handle (include/qsv.h).
pub struct DeviceState {
    handle: *mut ffi::QsvState,
    num_qubits: usize,
    _not_sync: std::marker::PhantomData<std::cell::Cell<()>>,
}

unsafe impl Send for DeviceState {}

impl DeviceState {
    /// `simulate`/`measure_all`/`get_state` are infallible in quantr (no `Result`), so FFI failures panic
    /// with a QuantrError-formatted message (SURVEY.md 8b "Error conventions").
    fn check(&self, code: c_int, what: &str) {
        if code != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::qsv_last_error(self.handle)) }.to_string_lossy().into_owned();
            panic!("\x1b[91m[Quantr Error] {what} failed in the device engine (code {code}): {msg}\x1b[0m ");
        }
    }

    /// One GPU, or - with `QSV_DEVICES=0,1,..` (a power of two of devices) - one handle over several GPUs of this
    /// process: the library shards the register itself (`qsv_create_multi`), `Circuit::simulate` never sees a device.
    /// Registers too small to shard (fewer than four qubits per device) stay on the first device.
    pub fn new(num_qubits: usize) -> DeviceState {
        let mut handle = std::ptr::null_mut();
        let mut devices: Vec<c_int> = std::env::var("QSV_DEVICES").ok()
            .map(|v| v.split(',').filter_map(|x| x.trim().parse().ok()).collect()).unwrap_or_default();
        while devices.len() > 1 && num_qubits < 4 + (usize::BITS - 1 - devices.len().leading_zeros()) as usize {
            devices.truncate(devices.len() / 2);
        }
        let code = if devices.len() > 1 {
            unsafe { ffi::qsv_create_multi(&mut handle, num_qubits as u32, devices.as_ptr(), devices.len() as c_int) }
        } else {
            unsafe { ffi::qsv_create(&mut handle, num_qubits as u32, devices.first().copied().unwrap_or(0)) }
        };
        let s = DeviceState { handle, num_qubits, _not_sync: std::marker::PhantomData };
        s.check(code, "qsv_create");
        s
    }

    pub fn init_basis(&mut self, index: u64) {
        let c = unsafe { ffi::qsv_init_basis(self.handle, index) };
        self.check(c, "qsv_init_basis");
    }

    /// `Circuit::change_register` (src/circuit.rs:463-473): Complex64 is two f64, interleaved.
    pub fn upload(&mut self, amplitudes: &[Complex64]) {
        let c = unsafe { ffi::qsv_upload(self.handle, amplitudes.as_ptr() as *const f64, 0, amplitudes.len() as u64) };
        self.check(c, "qsv_upload");
    }

    /// `SimulatedCircuit::get_state` / `take_state` (src/simulated_circuit.rs:158-160,185-187).
    pub fn download(&mut self) -> Vec<Complex64> {
        let mut out = vec![Complex64::new(0.0, 0.0); 1usize << self.num_qubits];
        let c = unsafe { ffi::qsv_download(self.handle, out.as_mut_ptr() as *mut f64, 0, out.len() as u64) };
        self.check(c, "qsv_download");
        out
    }

    pub fn apply(&mut self, ops: &EncodedOps) -> ffi::QsvStats {
        let mut stats = ffi::QsvStats::default();
        let c = unsafe { ffi::qsv_apply(self.handle, ops.ops.as_ptr(), ops.ops.len(), &mut stats) };
        self.check(c, "qsv_apply");
        stats
    }

    /// One `fastrand::f64()` per shot, drawn here in shot order (src/circuit/states/super_positions.rs:334), so
    /// `fastrand::seed` keeps its meaning; `u64::MAX` = "failed to collapse" (`None`, :341).
    pub fn sample(&mut self, shots: usize) -> Vec<u64> {
        let uniforms: Vec<f64> = (0..shots).map(|_| fastrand::f64()).collect();
        let mut out = vec![0u64; shots];
        let c = unsafe { ffi::qsv_sample(self.handle, uniforms.as_ptr(), shots as u64, out.as_mut_ptr()) };
        self.check(c, "qsv_sample");
        out
    }
}

impl DeviceState {
    /// Arbitrary canonical indices (registers too large for a host `Vec`, SURVEY.md 7.2 hard part 4).
    pub fn gather(&mut self, indices: &[u64]) -> Vec<Complex64> {
        let mut out = vec![Complex64::new(0.0, 0.0); indices.len()];
        let c = unsafe { ffi::qsv_gather(self.handle, indices.as_ptr(), indices.len() as u64, out.as_mut_ptr() as *mut f64) };
        self.check(c, "qsv_gather");
        out
    }

    pub fn norm_sqr(&mut self) -> f64 {
        let mut out = 0f64;
        let c = unsafe { ffi::qsv_norm_sqr(self.handle, &mut out) };
        self.check(c, "qsv_norm_sqr");
        out
    }
}

impl Drop for DeviceState {
    fn drop(&mut self) {
        unsafe { ffi::qsv_destroy(self.handle) };
    }
}

/// Custom gates on this many wires or more are passed as compact columns (include/qsv.h, qsv_op.iparam = 1).
pub const COMPACT_CUSTOM_WIRES: usize = 11;

/// `qsv_op[]` plus the buffers it borrows from for the duration of `qsv_apply`.
pub struct EncodedOps {
    pub ops: Vec<ffi::QsvOp>,
    controls: Vec<Vec<u32>>,
    matrices: Vec<Vec<f64>>,
    masks: Vec<Vec<u8>>,
}

/// What the encoder needs from one entry of `circuit_gates`; inside quantr this is a `match` over `Gate`
/// next to `Gate::linker` (src/circuit/gate.rs:140-168).
pub struct GateView<'a> {
    pub kind: u32, // 0 = Id, else QSV_GATE_*
    pub param: f64,
    pub iparam: i32,
    pub controls: &'a [usize],
    /// Custom only: the closure evaluated on basis sub-state `s` of [controls..., target] (MSB = first control):
    /// `None` = untouched, `Some(column)` = the 2^k amplitudes of its image.
    pub custom: Option<&'a dyn Fn(usize) -> Option<Vec<Complex64>>>,
}

pub fn encode(gates: &[GateView], num_qubits: usize) -> EncodedOps {
    let mut enc = EncodedOps { ops: Vec::new(), controls: Vec::new(), matrices: Vec::new(), masks: Vec::new() };
    for (counter, g) in gates.iter().enumerate() {
        if g.kind == 0 {
            continue; // Identity is skipped, src/circuit/simulation.rs:38-41
        }
        enc.controls.push(g.controls.iter().map(|&c| c as u32).collect());
        let ctrl = enc.controls.last().unwrap();
        let mut op = ffi::QsvOp {
            kind: g.kind,
            target: (counter % num_qubits) as u32, // src/circuit/simulation.rs:43
            n_controls: ctrl.len() as u32,
            reserved: 0,
            controls: if ctrl.is_empty() { std::ptr::null() } else { ctrl.as_ptr() },
            param: g.param,
            iparam: g.iparam,
            reserved2: 0,
            matrix: std::ptr::null(),
            none_mask: std::ptr::null(),
        };
        if let Some(closure) = g.custom {
            let k = g.controls.len() + 1;
            let dim = 1usize << k;
            if k >= COMPACT_CUSTOM_WIRES {
                // wide (multi-controlled) Custom gates: one 2^k column per sub-state the closure answers for,
                // ascending (qsv_op.iparam = 1); the engine lowers them to controlled ops of the fused pass
                let mut cols: Vec<f64> = Vec::new();
                let mut none = vec![1u8; dim];
                for s in 0..dim {
                    if let Some(column) = closure(s) {
                        none[s] = 0;
                        for t in 0..dim {
                            cols.push(column[t].re);
                            cols.push(column[t].im);
                        }
                    }
                }
                enc.matrices.push(cols);
                enc.masks.push(none);
                op.iparam = 1;
                op.matrix = enc.matrices.last().unwrap().as_ptr();
                op.none_mask = enc.masks.last().unwrap().as_ptr();
                enc.ops.push(op);
                continue;
            }
            let mut m = vec![0f64; 2 * dim * dim];
            let mut none = vec![0u8; dim];
            for s in 0..dim {
                match closure(s) {
                    None => none[s] = 1,
                    Some(column) => {
                        for t in 0..dim {
                            m[(t * dim + s) * 2] = column[t].re;
                            m[(t * dim + s) * 2 + 1] = column[t].im;
                        }
                    }
                }
            }
            enc.matrices.push(m);
            enc.masks.push(none);
            op.matrix = enc.matrices.last().unwrap().as_ptr();
            op.none_mask = enc.masks.last().unwrap().as_ptr();
        }
        enc.ops.push(op);
    }
    enc
}

#[allow(dead_code)]
fn _unused(_: *mut c_void) {}
