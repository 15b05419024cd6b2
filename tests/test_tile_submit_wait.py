"""isaac_ext_submit_build_templates / isaac_ext_wait (async submit + wait on a ticket): the submitted call gives the
result of the blocking call, one call in flight per context, tickets are checked."""
import os

import numpy as np
import pytest

import oracle_lib
from common_build import build_workload
from isaac_aligner_b200.batch import Tls, TemplateOptions
from isaac_aligner_b200.types import Config

pytestmark = pytest.mark.gpu


def test_submit_wait_build_templates():
    from isaac_aligner_b200 import capi
    genome, sim, reads, mb = build_workload(n_pairs=2000, L=100, seed=31)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    tls, options = Tls.make(), TemplateOptions.make(clip_semialigned=True)
    want = ctx.build_templates(mb, tls, options)
    ticket = ctx.submit_build_templates(mb, tls, options)
    with pytest.raises(capi.ExtError) as second:                         # one call in flight per context
        ctx.submit_build_templates(mb, tls, options)
    assert second.value.code == 4
    with pytest.raises(capi.ExtError):                                   # not the ticket that is in flight
        ctx.wait_templates(ticket + 1)
    for blocking in (lambda: ctx.set_reads(reads), lambda: ctx.build_templates(mb, tls, options),
                     lambda: ctx.determine_template_length(mb), lambda: ctx.trim_low_quality_ends(10)):
        with pytest.raises(capi.ExtError) as refused:                    # the context belongs to the submitted call until the wait
            blocking()
        assert refused.value.code == 4
    ctx.prefetch_reads(reads)                                            # staging the next tile is the one thing allowed meanwhile
    host_work = int(np.sort(np.arange(200_000)[::-1]).sum())             # the caller's thread is free meanwhile
    got = ctx.wait_templates(ticket)
    assert host_work > 0
    for name in want.templates.dtype.names:
        assert np.array_equal(got.templates[name], want.templates[name]), name
    for name in want.fragments.dtype.names:
        if name not in ("logProbability", "cigarOffset"):
            assert np.array_equal(got.fragments[name], want.fragments[name]), name
    assert np.array_equal(got.fragments["logProbability"].view(np.uint64), want.fragments["logProbability"].view(np.uint64))
    for i in range(len(want.fragments)):
        assert np.array_equal(got.cigar(i), want.cigar(i)), i
    with pytest.raises(capi.ExtError):                                   # the ticket is spent
        ctx.wait_templates(ticket)
    again = ctx.wait_templates(ctx.submit_build_templates(mb, tls, options))
    assert np.array_equal(again.templates["alignmentScore"], want.templates["alignmentScore"])
    ctx.submit_build_templates(mb, tls, options)                         # destroy joins a call nobody waited for
    ctx.close()
