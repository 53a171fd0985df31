//! golden_dump — runs the UNMODIFIED reference crate (Sin-tel/fft-convolution) on seeded inputs and
//! writes what it produced, so that the CPU oracle and the CUDA path of fft_convolution_b200 can be
//! held to reference-made vectors (tests/test_reference_golden.py).
//!
//! A case is: one convolver (`kind` + constructor arguments), a cyclic list of call sizes, and a
//! list of `update` / `reset` calls placed before given process() calls.  That is enough to replay
//! every case of tests/golden/make_golden.py and every scenario of the reference's own tests
//! (/root/reference/src/tests.rs:18-257, src/fft_convolver.rs:309-321, :528-540,
//! src/crossfade_convolver.rs:107-124) one convolver at a time.
//!
//! File format (little endian), one file per case, `ref_<name>.bin`:
//!   magic  b"FCBREF1\0"
//!   u32    number of entries
//!   entry: u32 name length, name bytes (utf-8), u32 dtype (0 = f32, 1 = utf-8 text), u64 count, payload
//! Text entry "meta" holds `key=value` pairs separated by ';' (see `Case::meta`).  f32 entries:
//! "h0", "h1", ... (impulse responses), "x" (input), "y" (the reference's output).
//!
//! Inputs are the synthetic data of SURVEY.md §8(d) (splitmix64 hash; white noise, random-decay
//! impulse responses) or the reference tests' f32 sinusoids (src/tests.rs:9-16); they are written
//! into the file, so the replay never has to regenerate them.

use std::fs::File;
use std::io::{BufWriter, Write};
use std::path::{Path, PathBuf};

use fft_convolution::crossfade_convolver::CrossfadeConvolver;
use fft_convolution::fft_convolver::{FFTConvolver, TwoStageFFTConvolver};
use fft_convolution::Convolution;

// ---- synthetic data (the same functions as oracle/fftconv_oracle.c orc_gen_*) ----
fn mix64(v: u64) -> u64 {
    let mut z = v.wrapping_add(0x9E37_79B9_7F4A_7C15);
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
    z ^ (z >> 31)
}

fn gen_noise(channel: u64, first_sample: usize, n: usize) -> Vec<f32> {
    let seed_x: u64 = 0x5EED_0001;
    (0..n)
        .map(|i| {
            let r = mix64(seed_x.wrapping_add(channel << 32).wrapping_add((first_sample + i) as u64));
            let u = (r >> 40) as f32 / 16_777_216.0_f32;
            2.0_f32 * u - 1.0_f32
        })
        .collect()
}

fn gen_ir(channel: u64, update_index: u64, len: usize) -> Vec<f32> {
    let seed_h: u64 = 0x5EED_0002_u64.wrapping_add(update_index << 48);
    let mut v = vec![0.0_f64; len];
    let mut energy = 0.0_f64;
    for i in 0..len {
        let r = mix64(seed_h.wrapping_add(channel << 32).wrapping_add(i as u64));
        let u = (r >> 40) as f64 / 16_777_216.0_f64;
        v[i] = (2.0 * u - 1.0) * (-6.9078_f64 * i as f64 / len as f64).exp();
        energy += v[i] * v[i];
    }
    let s = if energy > 0.0 { 1.0 / energy.sqrt() } else { 0.0 };
    v.iter().map(|&a| (a * s) as f32).collect()
}

// src/tests.rs:9-16 (all arithmetic in f32)
fn generate_sinusoid(length: usize, frequency: f32, sample_rate: f32, gain: f32) -> Vec<f32> {
    let mut signal = vec![0.0_f32; length];
    for i in 0..length {
        signal[i] = gain * (2.0 * std::f32::consts::PI * frequency * i as f32 / sample_rate).sin();
    }
    signal
}

// ---- case description ----
#[derive(Clone, Copy, PartialEq)]
enum Kind {
    Uniform,       // FFTConvolver::init(h0, block, max_len)
    TwoStage,      // TwoStageFFTConvolver::init(h0, block, max_len)
    CrossfadeNew,  // CrossfadeConvolver::new(FFTConvolver::init(h0, block, max_len), xf_len, xf_buf, fade)
    CrossfadeInit, // <CrossfadeConvolver<FFTConvolver> as Convolution>::init(h0, block, max_len)
}

struct Case {
    name: String,
    kind: Kind,
    block: usize,
    max_len: usize,
    xf_len: usize,
    xf_buf: usize,
    fade: usize,
    sizes: Vec<usize>,            // call sizes, cyclic
    updates: Vec<(usize, usize)>, // (before process() call number, index into irs)
    resets: Vec<usize>,           // before process() call number
    irs: Vec<Vec<f32>>,           // irs[0] initialises the convolver
    x: Vec<f32>,
}

impl Case {
    fn meta(&self) -> String {
        let kind = match self.kind {
            Kind::Uniform => "uniform",
            Kind::TwoStage => "twostage",
            Kind::CrossfadeNew => "crossfade_new",
            Kind::CrossfadeInit => "crossfade_init",
        };
        let list = |v: &Vec<usize>| v.iter().map(|a| a.to_string()).collect::<Vec<_>>().join(",");
        let upd = self.updates.iter().map(|(c, i)| format!("{}:{}", c, i)).collect::<Vec<_>>().join(",");
        format!(
            "kind={};block={};max_len={};xf_len={};xf_buf={};fade={};sizes={};updates={};resets={};n_irs={}",
            kind, self.block, self.max_len, self.xf_len, self.xf_buf, self.fade, list(&self.sizes), upd,
            list(&self.resets), self.irs.len()
        )
    }
}

enum Conv {
    U(FFTConvolver),
    T(TwoStageFFTConvolver),
    X(CrossfadeConvolver<FFTConvolver>),
}

impl Conv {
    fn process(&mut self, i: &[f32], o: &mut [f32]) {
        match self {
            Conv::U(c) => c.process(i, o),
            Conv::T(c) => c.process(i, o),
            Conv::X(c) => c.process(i, o),
        }
    }
    fn update(&mut self, r: &[f32]) {
        match self {
            Conv::U(c) => c.update(r),
            Conv::T(c) => c.update(r), // todo!() in the reference: no case asks for it
            Conv::X(c) => c.update(r),
        }
    }
    fn reset(&mut self) {
        match self {
            Conv::U(c) => c.reset(),
            Conv::T(c) => c.reset(),
            Conv::X(c) => c.reset(), // todo!() in the reference: no case asks for it
        }
    }
}

fn run(case: &Case) -> Vec<f32> {
    let h0 = &case.irs[0];
    let mut conv = match case.kind {
        Kind::Uniform => Conv::U(FFTConvolver::init(h0, case.block, case.max_len)),
        Kind::TwoStage => Conv::T(TwoStageFFTConvolver::init(h0, case.block, case.max_len)),
        Kind::CrossfadeNew => Conv::X(CrossfadeConvolver::new(
            FFTConvolver::init(h0, case.block, case.max_len),
            case.xf_len,
            case.xf_buf,
            case.fade,
        )),
        Kind::CrossfadeInit => Conv::X(<CrossfadeConvolver<FFTConvolver> as Convolution>::init(h0, case.block, case.max_len)),
    };
    let mut y = vec![0.0_f32; case.x.len()];
    let (mut p, mut call) = (0usize, 0usize);
    while p < case.x.len() {
        for &(at, which) in &case.updates {
            if at == call {
                conv.update(&case.irs[which]);
            }
        }
        if case.resets.contains(&call) {
            conv.reset();
        }
        let n = case.sizes[call % case.sizes.len()].min(case.x.len() - p);
        let mut out = vec![0.0_f32; n];
        conv.process(&case.x[p..p + n], &mut out);
        y[p..p + n].copy_from_slice(&out);
        p += n;
        call += 1;
    }
    y
}

fn write_case(dir: &Path, case: &Case, y: &[f32]) -> std::io::Result<PathBuf> {
    let path = dir.join(format!("ref_{}.bin", case.name));
    let mut w = BufWriter::new(File::create(&path)?);
    w.write_all(b"FCBREF1\0")?;
    let n_entries = 3 + case.irs.len() as u32; // meta, x, y, h*
    w.write_all(&n_entries.to_le_bytes())?;
    let text = |w: &mut BufWriter<File>, name: &str, s: &str| -> std::io::Result<()> {
        w.write_all(&(name.len() as u32).to_le_bytes())?;
        w.write_all(name.as_bytes())?;
        w.write_all(&1u32.to_le_bytes())?;
        w.write_all(&(s.len() as u64).to_le_bytes())?;
        w.write_all(s.as_bytes())
    };
    text(&mut w, "meta", &case.meta())?;
    let floats = |w: &mut BufWriter<File>, name: &str, v: &[f32]| -> std::io::Result<()> {
        w.write_all(&(name.len() as u32).to_le_bytes())?;
        w.write_all(name.as_bytes())?;
        w.write_all(&0u32.to_le_bytes())?;
        w.write_all(&(v.len() as u64).to_le_bytes())?;
        for a in v {
            w.write_all(&a.to_le_bytes())?;
        }
        Ok(())
    };
    for (i, h) in case.irs.iter().enumerate() {
        floats(&mut w, &format!("h{}", i), h)?;
    }
    floats(&mut w, "x", &case.x)?;
    floats(&mut w, "y", y)?;
    w.flush()?;
    Ok(path)
}

fn plain(name: &str, kind: Kind, block: usize, max_len: usize, sizes: Vec<usize>, h: Vec<f32>, x: Vec<f32>) -> Case {
    Case {
        name: name.to_string(), kind, block, max_len, xf_len: 0, xf_buf: 0, fade: 0, sizes,
        updates: vec![], resets: vec![], irs: vec![h], x,
    }
}

fn cases() -> Vec<Case> {
    let mut v = Vec::new();
    // ---- the cases of tests/golden/make_golden.py (same seeds: channel 11 / 12 / 13) ----
    v.push(plain("uniform_b64_l1000", Kind::Uniform, 64, 1000, vec![64], gen_ir(11, 0, 1000), gen_noise(11, 0, 64 * 24)));
    v.push(plain("uniform_b256_l3000_ragged", Kind::Uniform, 256, 3000, vec![100, 256, 37, 300], gen_ir(11, 0, 3000),
                 gen_noise(11, 0, 256 * 10)));
    v.push(plain("uniform_b512_l5000", Kind::Uniform, 512, 5000, vec![512], gen_ir(11, 0, 5000), gen_noise(11, 0, 512 * 8)));
    v.push(plain("twostage_h64_l12000", Kind::TwoStage, 64, 12000, vec![64], gen_ir(12, 0, 12000), gen_noise(12, 0, 64 * 80)));
    v.push(Case {
        name: "crossfade_b64_l300_update6".to_string(), kind: Kind::CrossfadeNew, block: 64, max_len: 300,
        xf_len: 300, xf_buf: 64, fade: 200, sizes: vec![64], updates: vec![(6, 1)], resets: vec![],
        irs: vec![gen_ir(13, 0, 300), gen_ir(13, 1, 300)], x: gen_noise(13, 0, 64 * 24),
    });
    // ---- BASELINE.json shapes, one channel each, short runs ----
    // configs[0]: FFTConvolver mono, block 256, 48 000-tap IR
    v.push(plain("cfg0_uniform_b256_l48000", Kind::Uniform, 256, 48000, vec![256], gen_ir(0, 0, 48000), gen_noise(0, 0, 256 * 400)));
    // configs[1]: TwoStage head 128, 240 000-tap IR (the reference derives T = 8192): 3 tail periods + a bit
    v.push(plain("cfg1_twostage_h128_l240000", Kind::TwoStage, 128, 240000, vec![128], gen_ir(0, 0, 240000),
                 gen_noise(0, 0, 128 * 200)));
    // configs[2]: CrossfadeConvolver::init, block 512, 96 000-tap IR, update every 50 blocks
    v.push(Case {
        name: "cfg2_crossfade_init_b512_l96000".to_string(), kind: Kind::CrossfadeInit, block: 512, max_len: 96000,
        xf_len: 0, xf_buf: 0, fade: 0, sizes: vec![512], updates: vec![(50, 1), (100, 2), (150, 1), (200, 2)], resets: vec![],
        irs: vec![gen_ir(0, 0, 96000), gen_ir(0, 1, 96000), gen_ir(0, 2, 96000)], x: gen_noise(0, 0, 512 * 260),
    });
    // configs[3]: one of the 4096 channels, block 512, 2 s IR
    v.push(plain("cfg3_uniform_b512_l96000", Kind::Uniform, 512, 96000, vec![512], gen_ir(7, 0, 96000), gen_noise(7, 0, 512 * 220)));
    // ---- the reference's own test scenarios, one convolver per case (src/tests.rs) ----
    let sr = 44100.0_f32;
    {
        // fft_convolver_update_is_reset (:18-59): update(b) before block 8
        let (a, b) = (generate_sinusoid(512, 1000.0, sr, 1.0), generate_sinusoid(512, 2000.0, sr, 0.7));
        let x = generate_sinusoid(16 * 512, 1300.0, sr, 1.0);
        let mut c = plain("reftest_update_is_reset", Kind::Uniform, 512, 512, vec![512], a.clone(), x.clone());
        c.irs.push(b);
        c.updates.push((8, 1));
        v.push(c);
        // test_crossfade_convolver (:61-117): new(conv_a, 512, 512, 512), update(b) before block 8
        let mut c = plain("reftest_crossfade_convolver", Kind::CrossfadeNew, 512, 512, vec![512], a, x);
        c.xf_len = 512;
        c.xf_buf = 512;
        c.fade = 512;
        c.irs.push(generate_sinusoid(512, 2000.0, sr, 0.7));
        c.updates.push((8, 1));
        v.push(c);
    }
    {
        // block_size_equal (:119-146): internal block 64 and 128, fed 128-sample calls
        let h = generate_sinusoid(128, 1000.0, sr, 0.1);
        let x = generate_sinusoid(200 * 128, 1300.0, sr, 1.0);
        v.push(plain("reftest_block_size_equal_b64", Kind::Uniform, 64, 128, vec![128], h.clone(), x.clone()));
        v.push(plain("reftest_block_size_equal_b128", Kind::Uniform, 128, 128, vec![128], h, x));
    }
    {
        // twostage_equal (:148-175): uniform(32) vs two-stage(64), 12 000-tap sinusoid IR, 64-sample calls
        let h = generate_sinusoid(12000, 1000.0, sr, 0.1);
        let x = generate_sinusoid(300 * 64, 1300.0, sr, 1.0);
        v.push(plain("reftest_twostage_equal_uniform_b32", Kind::Uniform, 32, 12000, vec![64], h.clone(), x.clone()));
        v.push(plain("reftest_twostage_equal_twostage_h64", Kind::TwoStage, 64, 12000, vec![64], h, x));
    }
    {
        // reset_fftconvolver / reset_twostagefftconvolver (:177-257): run, reset(), run again
        let h = generate_sinusoid(12000, 1000.0, sr, 0.1);
        let x1 = generate_sinusoid(300 * 64, 1300.0, sr, 0.1);
        let mut x = x1.clone();
        x.extend_from_slice(&x1);
        let mut c = plain("reftest_reset_uniform_b64", Kind::Uniform, 64, 12000, vec![64], h.clone(), x.clone());
        c.resets.push(300);
        v.push(c);
        let mut c = plain("reftest_reset_twostage_h64", Kind::TwoStage, 64, 12000, vec![64], h, x);
        c.resets.push(300);
        v.push(c);
    }
    {
        // the three delta-IR known-answer tests (src/fft_convolver.rs:309-321, :528-540, src/crossfade_convolver.rs:107-124)
        let mut d = vec![0.0_f32; 1024];
        d[0] = 1.0;
        let ones = vec![1.0_f32; 1024];
        v.push(plain("reftest_passthrough_uniform", Kind::Uniform, 1024, 1024, vec![1024], d.clone(), ones.clone()));
        v.push(plain("reftest_passthrough_twostage", Kind::TwoStage, 1024, 1024, vec![1024], d.clone(), ones.clone()));
        let mut c = plain("reftest_passthrough_crossfade", Kind::CrossfadeNew, 1024, 1024, vec![1024], d, ones);
        c.xf_len = 1024;
        c.xf_buf = 1024;
        c.fade = 1024;
        v.push(c);
    }
    v
}

fn main() -> std::io::Result<()> {
    let dir = std::env::args().nth(1).unwrap_or_else(|| "../../tests/golden".to_string());
    let dir = PathBuf::from(dir);
    std::fs::create_dir_all(&dir)?;
    for case in cases() {
        let y = run(&case);
        let path = write_case(&dir, &case, &y)?;
        let rms = (y.iter().map(|&a| (a as f64) * (a as f64)).sum::<f64>() / y.len().max(1) as f64).sqrt();
        println!("{}: {} samples, output rms {:.6}", path.display(), y.len(), rms);
    }
    Ok(())
}
