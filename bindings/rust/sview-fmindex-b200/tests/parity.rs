//! Parity with the real reference crate (needs a Rust toolchain + a B200): same blob, same patterns,
//! `count` equal and sorted `locate` equal.  NOT RUN IN THE BUILD IMAGE.
use sview_fmindex::{blocks::Block3, build_config::{LookupTableConfig, SuffixArrayConfig}, text_encoders::EncodingTable, FmIndex, FmIndexBuilder};
use sview_fmindex_b200::GpuFmIndex;

#[test]
fn gpu_matches_reference_on_a_reference_built_blob() {
    let symbols: &[&[u8]] = &[b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"];
    let table = EncodingTable::from_symbols(symbols);
    let text: Vec<u8> = (0..200_000u32).map(|i| b"ACGT"[(i.wrapping_mul(2654435761) >> 30) as usize]).collect();
    let builder = FmIndexBuilder::<u32, Block3<u64>, EncodingTable>::new(text.len(), table.symbol_count(), table).unwrap()
        .set_lookup_table_config(LookupTableConfig::KmerSize(3)).unwrap()
        .set_suffix_array_config(SuffixArrayConfig::Compressed(2)).unwrap();
    let mut blob = vec![0u8; builder.blob_size()];
    builder.build(text.clone(), &mut blob).unwrap();
    let cpu = FmIndex::<u32, Block3<u64>, EncodingTable>::load(&blob).unwrap();
    let gpu = GpuFmIndex::<u32, Block3<u64>, EncodingTable>::load(&blob).unwrap();
    for start in (0..text.len() - 20).step_by(997) {
        let pat = &text[start..start + 20];
        assert_eq!(cpu.count(pat), gpu.count(pat));
        let (mut a, mut b) = (cpu.locate(pat), gpu.locate(pat));
        assert_eq!(a, b, "SA-row order must match too");
        a.sort(); b.sort();
        assert_eq!(a, b);
    }
}
