"""CPU oracle for the ray-render hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``neural_graph_mapping_b200``) may import from
here.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the CPU baseline -- never as the thing shipped.

* ``oracle.restatement`` -- a plain-PyTorch (CPU, fp32) restatement of the
  reference algorithm, each function citing the reference file:line it follows.
  Pinned against the *unmodified* reference (imported from ``/root/reference``
  through ``oracle.ref_loader``) by ``oracle/make_golden.py`` -> ``tests/golden``.
* ``oracle.ref_loader`` -- imports the reference's own Python with stub modules
  for its missing third-party packages.  Works only where ``/root/reference``
  exists (the build container), never on the GPU box.
* ``oracle.permuto`` -- CPU restatement of the permutohedral-lattice hash
  encoding (third-party ``permutohedral_encoding``; source not in the reference
  tree).  PARITY UNPINNED -- see the module header.
"""
