"""GPU tests written after the round's GPU minutes were spent: they have run on the CPU build of the library only
(tests/test_hostsim_pipeline.py, tests/test_live_cli.py run the same checks there), never yet on a B200.  They sit in a
file that sorts last so that, under `pytest -x`, everything that has been seen green on the hardware runs first."""
import pytest

import orc
import pipeline_checks as pc
from test_live_cli import _exe, check_argv_matrix

pytestmark = pytest.mark.gpu


def test_dormant_prefilter_front_ends(pkg, gpu_lib):
    """opts.prefilter = 1..4 (k1_demod_pre_kernel) against the oracle, stages and lines"""
    pc.check_prefilter(pkg, gpu_lib)


def test_lane_event_overflow_costs_bits_not_the_stream(pkg, gpu_lib):
    """the fuzzer's two overflow regimes (lane event buffer, event ring): lines stay the reference's / the stream goes on"""
    pc.check_lane_event_overflow(pkg, gpu_lib)


@pytest.mark.skipif(orc.ref_binary() is None, reason="compiled reference (oracle/_ref) not present")
def test_argv_matrix_matches_the_reference_binary(pkg, gpu_lib):
    """SURVEY 8(b), argv by argv against the reference binary on the GPU build of the host program, without the three
    argvs that decode at a decimation no GPU test has run yet (0, 9): the decimations the GPU suite covers are 1-4"""
    check_argv_matrix(_exe(pkg), real_stdin=True, skip=("-d abc", "-d 0", "-d 9 -s"))


def test_telegram_that_ends_after_a_gap_in_the_input(pkg, gpu_lib):
    """time chunks: a telegram that runs into dead air ends 291 k samples after its chunk's end; the worker's right halo
    goes on while wmb_pending_before() reports it in flight (tests/test_time_shard.py has the CPU-build twin)"""
    import fuzz_cases
    from test_time_shard import check_time_chunks
    c = fuzz_cases.time_chunk_case(882, 126)
    check_time_chunks(pkg, gpu_lib, fuzz_cases.build_capture(c), c["flags"], world=4, halo_m=1 << 18)
