"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (plain torch-CPU / numpy, fp32 with an optional fp64 mode) of sdfstudio's per-ray SDF
volume-rendering hot path.  Each function cites the reference file:line it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the *checker* / the CPU arm -- never as the thing shipped.  ``sdfstudio_b200/``
must not import it (tests/test_no_oracle_in_product.py enforces that).

Parity pinning: the reference holds **no** golden vectors / known-answer tests for this path.
The oracle is therefore pinned against outputs of the *real reference code* imported in the build container
(``oracle/ref_import.py``; fixtures minted by ``oracle/make_golden.py`` and committed under ``tests/golden/``).
The tcnn-layout grid (``hashgrid.tcnn_*``) restates tiny-cuda-nn's published grid conventions; tiny-cuda-nn is an
unpinned, un-vendored dependency (Dockerfile:96) so that one variant is "parity unpinned" (see DESIGN.md).
"""
