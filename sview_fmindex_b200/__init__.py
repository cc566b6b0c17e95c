"""sview_fmindex_b200 -- B200-native (sm_100a) batched FM-index search: the `load` / `count` / `locate` hot
path of baku4/sview-fmindex on the reference's own blob format.  See DESIGN.md and include/svfm.h."""
from .fmindex import (BuildError, EmptyPattern, EncodingTable, FmIndex, FmIndexBuilder, IndexType, InvalidFormat,
                      LoadError, LookupTableConfig, MismatchedBlobSize, SuffixArrayConfig, SvfmError, aligned_empty)

__all__ = ["BuildError", "EmptyPattern", "EncodingTable", "FmIndex", "FmIndexBuilder", "IndexType", "InvalidFormat",
           "LoadError", "LookupTableConfig", "MismatchedBlobSize", "SuffixArrayConfig", "SvfmError", "aligned_empty"]
