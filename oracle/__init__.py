"""CPU oracle for the SpliceDICE quant / pairwise hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.  ``splicedice_b200`` never
imports this package and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference *executed in the build container*
(``oracle/gen_golden.py`` imports ``/root/reference/splicedice`` and
``scipy.stats.fisher_exact`` and writes ``tests/golden/*.npz|json``); the
fixtures and the generating script are committed, ``tests/test_oracle_golden.py``
replays them without the reference present.

Modules
-------
oracle_np     numpy restatement of cluster build, exclusion sums, PS (f32/f64),
              IR/RSD and Benjamini-Hochberg.
build_ref     pip-installs the UNMODIFIED reference into the git-ignored
              ``oracle/_ref/`` (travels to the GPU box with the tree).
ref_harness   drives that reference in memory; the timed CPU arm of bench.py
              (one process, or closed row slabs over all host cores).
ref_port      loop-for-loop port of the reference's Python hot loops (same data
              structures, same cost profile) -- timed only when no reference
              tree is installed.
fisher_c      ctypes wrapper over ``fisher_oracle.c`` (binary128 restatement of
              scipy's two-sided ``fisher_exact``).
"""
