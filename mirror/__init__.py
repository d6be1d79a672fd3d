"""NOT part of the product: a Python stand-in for the R shim (infercnv_b200/r/infercnv_b200.R), used by the tests and the
examples so that they read like the reference's own step sequence.  `ops` carries the reference's function names and
signatures (hspike mirroring, log lines, region report files) on top of infercnv_b200.api; every number comes from the
library."""
