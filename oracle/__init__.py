"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatements of the reference's hot-path algorithms, used only as the
checker by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under
``slenderobjdet_b200/`` imports this package.

* ``oracle.dcn``    -- DCN v1/v2 forward + backward (plain C, ``dcn_oracle.c``)
* ``oracle.assign`` -- pairwise_iou, Matcher, TopKMatcher (numpy)
* ``oracle.losses`` -- focal, smooth-L1, IoU/GIoU losses with gradients (numpy, float64 inside)
"""
