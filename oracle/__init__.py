"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's lifting path.

Nothing under `veon_b200/` may import this package.  It is used by `tests/`,
by `__graft_entry__.smoke()` and by `bench.py`'s CPU-baseline legs as the
*checker* (and as the timed CPU baseline), never as the product path.

Parity status (see DESIGN.md "Oracle"):
  * prepare (`lift_oracle.prepare_v2`)  -- PINNED: checked against the
    reference's own `voxel_pooling_prepare_v2` executed from /root/reference
    (tests/golden/make_golden.py -> tests/golden/*.npz, *.json).
  * pool fwd/bwd (`pool_oracle.c`)      -- PINNED by the reference's
    known-answer test (bev_pool.py:145-176) and, on a GPU box, against the
    reference's own kernels compiled unmodified into oracle/_ref/.
  * tail (`lift_oracle.voxel_text_labels`) -- parity UNPINNED: the reference
    has no test or fixture for it and its modules need detectron2/open_clip to
    import; it restates san_in_veon_temporal.py:257-259,
    san_in_veon_entry_temporal.py:273-297 and veon_temporal.py:223-229,240.
  * depth-distribution producer (`lift_oracle.downsample_depth`,
    `lift_oracle.two_hot_depth`)        -- PINNED: checked against the output of
    the reference's own downsample_depth / get_two_hot_depth
    (view_transformer_raw.py:393-429) executed from /root/reference
    (tests/golden/make_golden_depth.py -> tests/golden/two_hot_depth.npz).
  * 2x2x2 max-downsample and its gradient: the checker is the reference's own
    expression evaluated by ATen (`view(...).amax(dim=(3,5,7))`,
    view_transformer_raw.py:549-553), in tests/test_neck_gpu.py.
"""
