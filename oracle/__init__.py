"""CPU oracle for the trajectory hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a CPU (torch fp32 / numpy f64) restatement of the
reference algorithm for the path named in BASELINE.json.  It exists so that the
CUDA path in ``distillation_trajectories_b200`` can be checked against it.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package never does; it fails
loudly when its CUDA library is missing.

Pinning status: the reference's own tests hold no golden vectors for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against the reference ITSELF:
``oracle/make_golden.py`` imports the unmodified reference from /root/reference in
the build container, runs it on seeded inputs and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this
restatement (bit-exact or 1e-6), and ``tests/test_oracle_vs_reference.py`` compares
live against /root/reference whenever that tree is present.
"""
