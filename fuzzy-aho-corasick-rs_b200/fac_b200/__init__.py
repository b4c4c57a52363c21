"""fac_b200 -- host-side mirror of the fuzzy-aho-corasick-rs search interface over libfacgpu.so."""
from .api import (DEFAULT_THRESHOLD, FuzzyAhoCorasick, FuzzyAhoCorasickBuilder, FuzzyLimits, FuzzyMatch,
                  FuzzyMatches, FuzzyPenalties, FuzzyReplacer, GpuBackend, HaystackTooLarge, Order, Overlap,
                  Pattern, Prefiltered, SearchError, SearchOptions, Segment, StreamMatch)

__all__ = ["DEFAULT_THRESHOLD", "FuzzyAhoCorasick", "FuzzyAhoCorasickBuilder", "FuzzyLimits", "FuzzyMatch",
           "FuzzyMatches", "FuzzyPenalties", "FuzzyReplacer", "GpuBackend", "HaystackTooLarge", "Order",
           "Overlap", "Pattern", "Prefiltered", "SearchError", "SearchOptions", "Segment", "StreamMatch"]
