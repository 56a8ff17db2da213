"""CPU oracle for the Classpose per-tile post-network path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``classpose_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline -- never as the product.

Provenance / pinning status (see DESIGN.md "Oracle"):

* ``oracle.classpose_ref`` restates functions the reference repository owns
  (``compute_class_masks``, ``remove_border_instances``,
  ``unaugment_class_tiles``).  PINNED: checked against golden vectors produced
  by executing the reference's own source (AST-lifted from ``/root/reference``
  by ``tests/golden/make_golden.py``) and against the 8 known-answer cases of
  the reference's ``tests/test_remove_border_instances.py``.
* ``oracle.dynamics`` / ``oracle.utils`` / ``oracle.transforms`` restate the
  third-party package the reference calls for all other arithmetic:
  ``cellpose==4.0.8`` (uv.lock:352-353), with ``fastremap==1.17.7`` and
  ``fill-voids==2.1.1``.  That package is absent from ``/root/reference``, from
  this image and from the GPU box, and the reference's tests hold no golden
  vector for it.  **PARITY UNPINNED** for these modules: they follow the
  published algorithm op-for-op using the same library calls (torch
  ``grid_sample``, ``scipy.ndimage``), anchored on the reference call sites
  (models.py:120,149,172; core.py:130,173,209,215,218).
"""
