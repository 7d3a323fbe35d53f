"""Generates tests/golden/synth_0.002_seed12345.hgr by IMPORTING THE REFERENCE GENERATOR
(/root/reference/circuit_generator.py) with random.seed(12345) -- only possible in the build container.
tests/test_datasets.py checks that eig_kl_algorithm_b200.datasets.write_synthetic reproduces it byte for byte."""
import importlib.util
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("circuit_generator", "/root/reference/circuit_generator.py")
cg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cg)
random.seed(12345)
cg.FastCircuitGenerator(0.002).write_to_file(os.path.join(HERE, "synth_0.002_seed12345.hgr"))
