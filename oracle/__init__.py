"""CPU oracle for the ferreus BBFMM / RBF-solve hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy + a small
C helper) of the reference algorithms in graphic-goose/ferreus_rbf_rs.  It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` — never by the
product package ``ferreus_rbf_rs_b200``.

Parity status: the reference is pure Rust and cannot be compiled here (no
cargo/rustc), and its own tests hold no golden vectors for the
FMM matvec, tree contents, interaction lists, M2L operators or FGMRES iterates
("parity unpinned" for those).  The oracle is pinned against (i) exact dense
direct summation, (ii) the reproducible known-answer tests the reference does
hold (PointOutsideTree{1}, monomial bases, make_spd Cholesky, DDM invariants),
re-expressed in ``tests/``.
"""
