"""The oracle -> reference pin, re-run on the GPU box (`-m gpu`).

The parity tests proper compare the CUDA path with the oracle; what pins the oracle to the reference (the committed
golden vectors, the compiled reference in oracle/_ref, the host libm its NCO restatement follows) lives in
test_oracle_golden.py / test_oracle_vs_ref.py, which are unmarked and so only run in the CPU suite.  This module runs a
fast subset of them under the gpu marker as well, so that the whole chain reference -> oracle -> CUDA is verified on the
very box (and against the very libm) the GPU results are produced on.  No CUDA call is made here."""
import pytest

import test_oracle_golden as G
import test_oracle_vs_ref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,rs", [("rx_2400", 2400.0), ("rx_1200", 1200.0)])
def test_pin_rx_pipeline_golden(oracle_lib, golden, name, rs):
    G.test_rx_pipeline_matches_reference_golden(oracle_lib, golden, name, rs)


def test_pin_tx_fir_bits_fft_golden(oracle_lib, golden):
    G.test_tx_matches_reference_golden(oracle_lib, golden)
    G.test_appendix_b_known_answers(oracle_lib, golden)
    G.test_fir256_matches_reference_golden(oracle_lib, golden)
    G.test_bit_stages_match_reference_golden(oracle_lib, golden)
    G.test_fft_matches_reference_golden(oracle_lib, golden)


@pytest.mark.parametrize("flavour,rs,esn0", [("2400", 2400.0, 12.0), ("1200", 1200.0, 20.0)])
def test_pin_rx_against_compiled_reference(oracle_lib, flavour, rs, esn0):
    R.test_rx_against_compiled_reference(oracle_lib, flavour, rs, esn0)


def test_pin_loop_functions_and_tx_lengths(oracle_lib):
    R.test_costas_loop_functions(oracle_lib)
    R.test_tx_any_length_against_compiled_reference(oracle_lib)


def test_pin_nco_restatement_against_this_box_libm(oracle_lib):
    R.test_glibc_sincos_restatement_matches_host_libm(oracle_lib)
