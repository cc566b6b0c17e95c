//! Writes golden blobs built by the REAL reference crate, for tests/test_rust_fixtures.py to pin the byte layout of
//! this repo's builders (oracle/fm_oracle.c and csrc/builder.cu) against.  Run once on a machine with cargo:
//!
//!     cargo run --release --example dump_fixtures -- ../../../tests/golden/rust_blobs
//!
//! and commit the files it writes.  One fixture per Vector width and Position width.  NOT RUN IN THE BUILD IMAGE
//! (no Rust toolchain there); until somebody runs it, blob byte parity with the Rust builder stays unpinned.
use std::{fs, io::Write, path::PathBuf};
use sview_fmindex::{blocks::{Block2, Block3, Block5}, build_config::{LookupTableConfig, SuffixArrayConfig},
                    text_encoders::EncodingTable, FmIndex, FmIndexBuilder, Position, Block};

fn text(n: usize, alphabet: &[u8]) -> Vec<u8> {
    // splitmix64 counter generator of sview_fmindex_b200/synth.py (seed 42), so the Python side can regenerate the text
    (0..n as u64).map(|i| {
        let mut x = 42u64.wrapping_add(0x9E3779B97F4A7C15u64.wrapping_mul(i + 1)).wrapping_add(0x9E3779B97F4A7C15);
        x = (x ^ (x >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        x = (x ^ (x >> 27)).wrapping_mul(0x94D049BB133111EB);
        x ^= x >> 31;
        alphabet[(((x >> 32) * alphabet.len() as u64) >> 32) as usize]
    }).collect()
}

fn dump<P: Position + std::fmt::Debug, B: Block>(dir: &PathBuf, name: &str, symbols: &[&[u8]], wildcard: bool, alphabet: &[u8],
                                                  n: usize, k: u32, r: u32, pos_bits: u32, planes: u32, vec_bits: u32) {
    let table = if wildcard { EncodingTable::from_symbols_with_wildcard(symbols) } else { EncodingTable::from_symbols(symbols) };
    let t = text(n, alphabet);
    let builder = FmIndexBuilder::<P, B, EncodingTable>::new(t.len(), table.symbol_count(), table).unwrap()
        .set_lookup_table_config(if k == 1 { LookupTableConfig::None } else { LookupTableConfig::KmerSize(k) }).unwrap()
        .set_suffix_array_config(if r == 1 { SuffixArrayConfig::Uncompressed } else { SuffixArrayConfig::Compressed(r) }).unwrap();
    let mut blob = vec![0u8; builder.blob_size()];
    builder.build(t.clone(), &mut blob).unwrap();
    let ix = FmIndex::<P, B, EncodingTable>::load(&blob).unwrap();
    let mut queries = Vec::new();
    for (qi, start) in (0..t.len().saturating_sub(12)).step_by(t.len() / 23 + 1).enumerate() {
        let pat = &t[start..start + 1 + qi % 11];
        queries.push(format!("{{\"pattern_hex\": \"{}\", \"count\": {:?}, \"locate\": {:?}}}",
                             pat.iter().map(|b| format!("{b:02x}")).collect::<String>(), ix.count(pat), ix.locate(pat)));
    }
    fs::write(dir.join(format!("{name}.bin")), &blob).unwrap();
    let mut f = fs::File::create(dir.join(format!("{name}.json"))).unwrap();
    write!(f, "{{\"name\": \"{name}\", \"text_len\": {n}, \"text_seed\": 42, \"alphabet\": \"{}\", \"symbols\": [{}], \"wildcard\": {wildcard}, \
               \"pos_bits\": {pos_bits}, \"planes\": {planes}, \"vec_bits\": {vec_bits}, \"kmer_size\": {k}, \"sampling_ratio\": {r}, \
               \"blob_file\": \"{name}.bin\", \"blob_len\": {}, \"queries\": [{}]}}\n",
           String::from_utf8_lossy(alphabet),
           symbols.iter().map(|s| format!("\"{}\"", String::from_utf8_lossy(s))).collect::<Vec<_>>().join(", "),
           blob.len(), queries.join(", ")).unwrap();
}

fn main() {
    let dir = PathBuf::from(std::env::args().nth(1).expect("output directory"));
    fs::create_dir_all(&dir).unwrap();
    let dna: &[&[u8]] = &[b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"];
    let dna4: &[&[u8]] = &[b"Aa", b"Cc", b"Gg", b"Tt"];
    let amino: Vec<Vec<u8>> = b"ACDEFGHIKLMNPQRSTVWY".iter().map(|c| vec![*c]).collect();
    let amino: Vec<&[u8]> = amino.iter().map(|v| v.as_slice()).collect();
    dump::<u32, Block3<u64>>(&dir, "u32_block3_u64", dna, false, b"ACGT", 70_001, 3, 2, 32, 3, 64);
    dump::<u32, Block2<u32>>(&dir, "u32_block2_u32", dna4, false, b"ACGT", 4_097, 2, 3, 32, 2, 32);
    dump::<u64, Block2<u128>>(&dir, "u64_block2_u128", dna4, false, b"ACGT", 12_800, 1, 1, 64, 2, 128);
    dump::<u64, Block3<u64>>(&dir, "u64_block3_u64", dna, false, b"ACGT", 30_000, 3, 4, 64, 3, 64);
    dump::<u32, Block5<u64>>(&dir, "u32_block5_u64_protein", &amino, true, b"ACDEFGHIKLMNPQRSTVWYX", 20_000, 3, 2, 32, 5, 64);
    dump::<u64, Block3<u32>>(&dir, "u64_block3_u32", dna, false, b"ACGTN", 9_999, 4, 16, 64, 3, 32);
}
