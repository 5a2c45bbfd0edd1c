"""CPU checks of the persistent sweep kernel's static schedule (mmmfe-st-dd_b200/csrc/sweep_plan.cpp): the host
builds the per-CTA entry lists, ring offsets and issue points, then replays them against the kernel's own rules
(ring ranges never overwritten before their consumer finished, barrier slots not re-armed early, every completion
counter reaches its target, no deadlock in the data-flow waits)."""
import pytest

import mmmfe_import


@pytest.mark.parametrize("nx,ny,leaf,sub,M,ncta", [
    (3, 3, 4, 256, 1, 296),        # default parameter.txt, cycle 0: the whole subdomain is one subtree
    (2, 2, 4, 256, 1, 296),
    (16, 16, 4, 64, 2, 296),
    (37, 23, 4, 64, 2, 296),       # ragged boxes
    (96, 96, 4, 64, 2, 296),
    (150, 130, 4, 64, 4, 148),
    (5, 7, 1, 2, 1, 8),            # tiny leaves, few CTAs
    (256, 256, 4, 64, 4, 296),     # BASELINE configs[3] fine subdomain, 4 right-hand sides
    (512, 512, 4, 64, 2, 296),     # BASELINE configs[1] coarse subdomain
])
def test_schedule_replays_without_hazards(nx, ny, leaf, sub, M, ncta):
    pkg = mmmfe_import.load()
    rc, stats, msg = pkg.sweep_plan_check(nx, ny, leaf, sub, M, ncta, 115000, 3)
    assert rc == 0, msg
    assert stats["entries"] >= 2 and stats["ring_doubles"] >= 2 * stats["max_entry_doubles"]
    assert stats["smem_bytes"] <= 115000


def test_oversized_subtrees_are_refused_not_truncated():
    pkg = mmmfe_import.load()
    rc, _, msg = pkg.sweep_plan_check(64, 64, 4, 256, 1, 296, 115000, 1)
    assert rc == -3 and "too small" in msg


def test_trees_without_subtrees_are_refused():
    pkg = mmmfe_import.load()
    rc, _, msg = pkg.sweep_plan_check(32, 32, 4, 0, 1, 296, 115000, 1)
    assert rc == -2 and "subtree" in msg
