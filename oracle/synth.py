"""Synthetic weights / lines recipe shared by the oracle scripts and the tests (re-export; the generator is plain
data generation and lives in ocrfa_b200/synth.py so that bench.py's GPU arm does not have to import oracle/)."""
from ocrfa_b200.synth import bn_name_for, resnet_conv_specs, synth_lines, synth_state_dict  # noqa: F401
