"""Synthetic inputs for tests, bench.py and the smoke test: seeded windows of the BASELINE configurations
(SURVEY.md 8d), image pairs for KLT, PnP problems.  Test-input generators only -- not part of the product
package pvio_b200 (which holds the CUDA library and the host mirror of the reference interface)."""
