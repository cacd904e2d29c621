//! The reference's `benches/rand_benchmarks.rs:7-121` re-targeted at the batch entry points of the shim (the GPU is
//! built for batches; one pair per call measures launch latency).  Same five groups, same seed (1234), same equality
//! assertions between the crate's names (`rand_benchmarks.rs:17-21, 45-46, 65-67, 88-90, 113-114`), but every group
//! works on BATCH random pairs / haystacks and reports throughput in pairs (haystacks) per second.
//!
//! NOT COMPILED IN THIS REPOSITORY (neither the build image nor the GPU box has cargo: profiles/r02_probe_gpu_box.txt);
//! `python bench.py` measures the same workloads through the same C ABI.  To run it where a Rust toolchain exists:
//!   LD_LIBRARY_PATH=../../../triple_accel_b200 cargo bench      (dev-dependencies: criterion 0.3, rand 0.7.3)
use criterion::*;
use rand::prelude::*;
use triple_accel::hamming::*;
use triple_accel::levenshtein::*;
use triple_accel::*;

const BATCH: usize = 100_000;

fn rand_str<R: Rng>(len: usize, rng: &mut R) -> Vec<u8> {
    (0..len).map(|_| rng.gen_range(33u8, 127u8)).collect()
}

/// b = a with up to k random substitutions / insertions / deletions (model: rand_levenshtein_mutate)
fn mutate<R: Rng>(a: &[u8], k: u32, rng: &mut R) -> Vec<u8> {
    let mut b = a.to_vec();
    for _ in 0..rng.gen_range(k / 2, k + 1) {
        match rng.gen_range(0, 3) {
            0 if !b.is_empty() => {
                let p = rng.gen_range(0, b.len());
                b[p] = 32u8;
            }
            1 => {
                let p = rng.gen_range(0, b.len() + 1);
                b.insert(p, rng.gen_range(33u8, 127u8));
            }
            _ if !b.is_empty() => {
                let p = rng.gen_range(0, b.len());
                b.remove(p);
            }
            _ => {}
        }
    }
    b
}

fn csr(strs: &[Vec<u8>]) -> (Vec<u8>, Vec<u64>) {
    let mut off = vec![0u64];
    let mut buf = Vec::new();
    for s in strs {
        buf.extend_from_slice(s);
        off.push(buf.len() as u64);
    }
    (buf, off)
}

fn bench_rand_hamming(c: &mut Criterion) {
    let mut rng = StdRng::seed_from_u64(1234);
    let mut group = c.benchmark_group("bench_rand_hamming_batch");
    for str_len in [10usize, 100, 1000].iter() {
        let a: Vec<Vec<u8>> = (0..BATCH).map(|_| rand_str(*str_len, &mut rng)).collect();
        let b: Vec<Vec<u8>> = a.iter().map(|s| { let mut t = s.clone(); for _ in 0..(*str_len / 10) { let p = rng.gen_range(0, t.len()); t[p] = 32u8; } t }).collect();
        let (ab, ao) = csr(&a);
        let (bb, bo) = csr(&b);
        let res = hamming_batch(&ab, &ao, &bb, &bo);
        assert!(res[0] == hamming_naive(&a[0], &b[0]) && res[0] == hamming_words_64(&a[0], &b[0])
                && res[0] == hamming_words_128(&a[0], &b[0]) && res[0] == hamming_simd_movemask(&a[0], &b[0])
                && res[0] == hamming_simd_parallel(&a[0], &b[0]));
        group.throughput(Throughput::Elements(BATCH as u64));
        group.bench_function(BenchmarkId::new("hamming_batch", *str_len), |bch| bch.iter(|| hamming_batch(&ab, &ao, &bb, &bo)));
    }
    group.finish();
}

fn bench_rand_levenshtein(c: &mut Criterion) {
    let mut rng = StdRng::seed_from_u64(1234);
    let mut group = c.benchmark_group("bench_rand_levenshtein_batch");
    for str_len in [10usize, 100, 1000].iter() {
        let k = (*str_len as u32) / 10;
        let a: Vec<Vec<u8>> = (0..BATCH).map(|_| rand_str(*str_len, &mut rng)).collect();
        let b: Vec<Vec<u8>> = a.iter().map(|s| mutate(s, k, &mut rng)).collect();
        let (ab, ao) = csr(&a);
        let (bb, bo) = csr(&b);
        let exact = levenshtein_exp_batch(&ab, &ao, &bb, &bo, LEVENSHTEIN_COSTS);
        let bounded = levenshtein_simd_k_batch(&ab, &ao, &bb, &bo, k, LEVENSHTEIN_COSTS);
        for i in 0..16 {
            // rand_benchmarks.rs:65-67, 88-90
            assert!(exact[i] == levenshtein_naive(&a[i], &b[i]) && exact[i] == levenshtein(&a[i], &b[i]));
            assert!(bounded[i] == levenshtein_naive_k_with_opts(&a[i], &b[i], k, false, LEVENSHTEIN_COSTS).map(|x| x.0));
        }
        group.throughput(Throughput::Elements(BATCH as u64));
        group.bench_function(BenchmarkId::new("levenshtein_exp_batch", *str_len), |bch| bch.iter(|| levenshtein_exp_batch(&ab, &ao, &bb, &bo, LEVENSHTEIN_COSTS)));
        group.bench_function(BenchmarkId::new("levenshtein_simd_k_batch", *str_len), |bch| bch.iter(|| levenshtein_simd_k_batch(&ab, &ao, &bb, &bo, k, LEVENSHTEIN_COSTS)));
    }
    group.finish();
}

fn bench_rand_levenshtein_search(c: &mut Criterion) {
    let mut rng = StdRng::seed_from_u64(1234);
    let mut group = c.benchmark_group("bench_rand_levenshtein_search_batch");
    for str_len in [100usize, 1000].iter() {
        let needle_len = *str_len / 10;
        let k = (*str_len as u32) / 100;
        let needle = rand_str(needle_len, &mut rng);
        let hays: Vec<Vec<u8>> = (0..BATCH / 10).map(|_| {
            let mut h = rand_str(*str_len, &mut rng);
            if rng.gen_range(0, 20) == 0 {
                let m = mutate(&needle, k, &mut rng);
                let p = rng.gen_range(0, h.len() - m.len());
                h[p..p + m.len()].copy_from_slice(&m);
            }
            h
        }).collect();
        let (hb, ho) = csr(&hays);
        let (m, mo) = levenshtein_search_batch(&needle, &hb, &ho, k, SearchType::All, LEVENSHTEIN_COSTS, false);
        // rand_benchmarks.rs:113-114 on the first haystack
        let first: Vec<Match> = levenshtein_search_naive_with_opts(&needle, &hays[0], k, SearchType::All, LEVENSHTEIN_COSTS, false).collect();
        assert!(first[..] == m[mo[0] as usize..mo[1] as usize]);
        group.throughput(Throughput::Elements(hays.len() as u64));
        group.bench_function(BenchmarkId::new("levenshtein_search_batch", *str_len), |bch| bch.iter(|| levenshtein_search_batch(&needle, &hb, &ho, k, SearchType::All, LEVENSHTEIN_COSTS, false)));
    }
    group.finish();
}

criterion_group!(bench_rand, bench_rand_hamming, bench_rand_levenshtein, bench_rand_levenshtein_search);
criterion_main!(bench_rand);
