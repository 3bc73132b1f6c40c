"""CPU oracle for the PointPillars input path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``3d-object-detection_b200``) never imports it and has no CPU fallback.

Contents (every function cites the reference file:line it follows):

* ``native``     -- ctypes wrapper over ``pp_oracle.c`` (create_pillars, iou, make_ious).
* ``ref``        -- loader for ``oracle/_ref/pillars*.so``: the reference's own
                    ``data/pillars.cpp`` compiled unmodified against the Boost stand-in in
                    ``oracle/boost_shim`` (Boost is absent from the image).
* ``glue``       -- numpy/torch restatement of ``data/dataset.py:88-106``.
* ``targets``    -- numpy restatement of ``utils/box_utils.py`` (anchors, boxes_to_image_space,
                    make_target, create_target).
* ``pfn``        -- fp64 restatement of ``model/model.py:13-62`` (PPFeatureNet, PPScatter).
* ``exact_iou``  -- exact rational (fractions.Fraction) convex clip used to pin the IoU.

Parity pinning status: see the header of ``pp_oracle.c`` and DESIGN.md ("Oracle").
"""
