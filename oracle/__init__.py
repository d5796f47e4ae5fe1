"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference hot path (kirstenmaas/nerf-for-angiography,
``nerf/run_nerf_acc.py:284-307``) used as the parity checker and as the timed CPU
baseline.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
package ``nerf_for_angiography_b200`` never does.

Pinning status (see DESIGN.md section "Oracle"):
  * geometry, CPPN, get_predictions, acc_render_volume_density, midpoint positions:
    PINNED against the reference's own Python code imported unmodified in the build
    container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
  * nerfacc ray marching / visibility / occupancy grid (third-party, absent, unpinned
    by the reference): PARITY UNPINNED -- restated from the nerfacc 0.3.5 sources.
"""
