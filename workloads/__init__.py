"""Synthetic workloads of BASELINE.json (C1..C5) and the small scenes of the parity tests: host-side scene
descriptions (numpy) driven by bench.py, the tests and the experiment scripts.  Not part of the product package --
fountain_b200/ holds only the path behind the C ABI."""
