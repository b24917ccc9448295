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
  * tail (`lift_oracle.class_groups`, `lift_oracle.voxel_text_labels`) -- PINNED
    for the prompt groups, the logits and the class merge: checked against the
    reference's own `_add_vocabulary_nuscenes`, `semantic_inference_3d` and
    `_merge_classes_prob` (san_in_veon_entry_temporal.py:243-262,273-297,
    san_in_veon_temporal.py:257-259) executed from /root/reference with
    detectron2 / open_clip stubbed (tests/golden/make_golden_tail.py ->
    tests/golden/tail_reference.npz).  The label rule (veon_temporal.py:223-229,
    240) is inline code of `simple_test`, not callable on its own: the fixture
    applies those lines literally (restated, not executed from the reference).
    The text encoder (open_clip, un-vendored, no version pin) is out of scope:
    the classifier weight is a synthetic constant.
  * trilinear up-sampling (`lift_oracle.trilinear_upsample`) -- PINNED against
    torch.nn.functional.interpolate, the function the reference calls
    (san_in_veon_temporal.py:196-207), at test time.
  * depth-distribution producer (`lift_oracle.downsample_depth`,
    `lift_oracle.two_hot_depth`)        -- PINNED: checked against the output of
    the reference's own downsample_depth / get_two_hot_depth
    (view_transformer_raw.py:393-429) executed from /root/reference
    (tests/golden/make_golden_depth.py -> tests/golden/two_hot_depth.npz).
  * 2x2x2 max-downsample and its gradient: the checker is the reference's own
    expression evaluated by ATen (`view(...).amax(dim=(3,5,7))`,
    view_transformer_raw.py:549-553), in tests/test_neck_gpu.py.
"""
