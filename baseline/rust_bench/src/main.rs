//! cargo run --release -- <uniform|zipf> <bytes>
//! Prints one JSON line: round-trip GB/s of input for huff_coding::prelude::{compress, decompress}.
//! The generators are the integer-only ones of huff_encoding_b200/datagen.py (splitmix64), so the bytes are identical.
use huff_coding::prelude::{compress, decompress};
use std::time::Instant;

fn splitmix64(mut x: u64) -> u64 {
    x = x.wrapping_add(0x9E3779B97F4A7C15);
    x = (x ^ (x >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    x = (x ^ (x >> 27)).wrapping_mul(0x94D049BB133111EB);
    x ^ (x >> 31)
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let kind = args.get(1).map(|s| s.as_str()).unwrap_or("uniform");
    let n: usize = args.get(2).and_then(|s| s.parse().ok()).unwrap_or(64 << 20);
    let seed: u64 = 0x5EED0000 + if kind == "zipf" { 2 } else { 1 };
    let data: Vec<u8> = if kind == "zipf" {
        // P(k) ~ k^-1.2, k = 1..256, u32 thresholds as in datagen.zipf_table()
        let p: Vec<f64> = (1..=256).map(|k| (k as f64).powf(-1.2)).collect();
        let s: f64 = p.iter().sum();
        let w: Vec<u64> = p.iter().map(|x| ((x / s) * 4294967296.0).round().max(1.0) as u64).collect();
        let tot: u64 = w.iter().sum();
        let mut thr = Vec::with_capacity(256);
        let mut c = 0u64;
        for x in &w { c += x; thr.push(c * 4294967296 / tot); }
        (0..n as u64).map(|i| { let u = splitmix64(seed.wrapping_add(i)) >> 32; thr.partition_point(|&t| t <= u) as u8 }).collect()
    } else {
        (0..n as u64).map(|i| (splitmix64(seed.wrapping_add(i)) & 0xFF) as u8).collect()
    };
    let t0 = Instant::now();
    let cd = compress(&data);
    let back = decompress(&cd);
    let dt = t0.elapsed().as_secs_f64();
    assert_eq!(back, data);
    println!("{{\"impl\": \"reference\", \"kind\": \"reference\", \"workload\": \"{}\", \"bytes\": {}, \"seconds\": {:.6}, \"value\": {:.6}, \"unit\": \"GB/s\", \"cores\": 1}}",
             kind, n, dt, n as f64 / dt / 1e9);
}
