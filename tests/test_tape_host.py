"""CPU tests of the tape host logic (no kernels run): the recorded programs of the fusion path have the structure the
native executor validates, every parameter the reference trains in that branch is registered exactly once, the T = 1
specialisation drops exactly the single-key attentions, and the C-side validator rejects malformed programs."""
import ctypes as C
from argparse import Namespace

import pytest

ARGS = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                 aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))


@pytest.fixture(scope="module")
def model():
    import mil_b200
    return mil_b200.get_model(ARGS)


def _kinds(t):
    from mil_b200 import _lib as L
    names = {L.OP_LINEAR: "linear", L.OP_ATTENTION: "attention", L.OP_LAYERNORM: "layernorm", L.OP_ADD: "add"}
    out = {}
    for o in t.ops:
        out[names[o[0]]] = out.get(names[o[0]], 0) + 1
    return out


def test_fusion_program_structure(model):
    t = model._fusion_tape(False)
    k = _kinds(t)
    # per TwoWayTransformer call: 2 blocks x (3 attentions x 4 linears + 2 MLP linears) + final attention 4 = 32 linears,
    # 7 attention cores, 9 LayerNorms, 2 shared keys+pe sums; two calls + fc_pathology, fc_CI2CT, fc_CI2Pth
    assert k == {"linear": 67, "attention": 14, "layernorm": 18, "add": 4}
    assert len(t.inputs) == 5 and len(t.buffers) == 1 and len(t.outputs) == 4
    rows = {"T": 3, "Nc": 160, "Np": 1000}
    assert t.buffers[0][0](rows) == 2 * 3 + 160 + 1000                       # aggregator.py:173 bag length
    offs = [fn(rows) for _, _, fn in t.outputs]
    assert offs == [0, 3, 163, 166]                                           # x_CT2CI | x_CI2CT | x_Pth2CI | x_CI2Pth
    # the program's parameters = fc_pathology, fc_CI2CT, fc_CI2Pth and TwoWayTransformer_Both, each exactly once
    want = {id(p) for n, p in model.named_parameters()
            if n.startswith(("TwoWayTransformer_Both.", "fc_pathology.", "fc_CI2CT.", "fc_CI2Pth."))}
    got = [id(p) for p in t.params]
    assert len(got) == len(set(got)) and set(got) == want
    c = t._freeze()
    assert c["total"] == sum(p.numel() for p in t.params) and all(o % 8 == 0 for o in c["offsets"])


def test_single_token_program_drops_single_key_attentions(model):
    t1, tn = model._fusion_tape(True), model._fusion_tape(False)
    k1 = _kinds(t1)
    # per call: 2 self-attentions and 2 image->token attentions lose their core and their q/k projections
    assert k1["attention"] == 14 - 2 * 4 and k1["linear"] == 67 - 2 * 4 * 2 and k1["layernorm"] == 18
    # ... but their q/k projection parameters stay registered (they receive exactly-zero gradients)
    assert {id(p) for p in t1.params} == {id(p) for p in tn.params}


def test_native_validator_rejects_malformed_programs(model):
    import mil_b200
    from mil_b200 import _lib as L
    lib = mil_b200.lib()
    t = model.TwoWayTransformer_Both._tape()
    c = t._freeze()
    rows = {"N": 100, "T": 2}
    slots = t._slots(rows)
    assert lib.milb200_tape_arena_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], L.F32, None) > 100 * 512 * 4
    n = c["n_slots"]
    ext = (C.c_void_p * n)()
    dummy = C.c_void_p(256)
    # (1) missing pointers
    rc = lib.milb200_tape_forward(c["ops"], c["n_ops"], slots, n, c["params"], c["n_params"], ext, None, None, None, 0, None,
                                  0, L.F32, None, None)
    assert rc != 0 and b"null" in lib.milb200_last_error()
    # (2) a shape that contradicts the weights: shrink the column count of one linear's input slot
    bad = type(slots).from_buffer_copy(slots)            # _slots() caches per shape: never edit its result
    lin = next(o for o in t.ops if o[0] == L.OP_LINEAR)
    bad[lin[1]].cols = 511
    rc = lib.milb200_tape_forward(c["ops"], c["n_ops"], bad, n, c["params"], c["n_params"], ext, dummy, dummy, dummy, 1 << 30,
                                  dummy, 1 << 30, L.F32, None, None)
    assert rc != 0 and b"shape mismatch" in lib.milb200_last_error()
    # (3) unknown op kind
    ops = (L.TapeOp * c["n_ops"])(*c["ops"])
    ops[0].kind = 99
    rc = lib.milb200_tape_forward(ops, c["n_ops"], slots, n, c["params"], c["n_params"], ext, dummy, dummy, dummy, 1 << 30,
                                  dummy, 1 << 30, L.F32, None, None)
    assert rc != 0 and b"unknown kind" in lib.milb200_last_error()
    # (4) only two lanes exist
    ops = (L.TapeOp * c["n_ops"])(*c["ops"])
    ops[3].lane = 2
    rc = lib.milb200_tape_forward(ops, c["n_ops"], slots, n, c["params"], c["n_params"], ext, dummy, dummy, dummy, 1 << 30,
                                  dummy, 1 << 30, L.F32, None, None)
    assert rc != 0 and b"lane" in lib.milb200_last_error()


def test_fusion_program_puts_the_ct_branch_on_its_own_lane(model):
    """aggregator.py:160 (CT) and :168 (pathology) are independent until the bag is assembled: the CT branch's ops carry
    lane 1, everything else lane 0, and a second lane doubles the per-lane scratch the executor asks for."""
    import mil_b200
    from mil_b200 import _lib as L
    lib = mil_b200.lib()
    t = model._fusion_tape()
    lanes = [o[8] for o in t.ops]
    assert set(lanes) == {0, 1}
    ct_in = t.inputs[0]
    reach = {ct_in}
    for o in t.ops:                                   # everything computed from the CT tokens sits on lane 1
        if any(x in reach for x in o[1:4] if x >= 0):
            reach.add(o[4])
            assert o[8] == 1
    assert t.ops[0][8] == 0                           # fc_pathology
    c = t._freeze()
    rows = {"T": 1, "Nc": 160, "Np": 300}
    slots = t._slots(rows)
    two = lib.milb200_tape_workspace_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], L.BF16, 0, None)
    ops1 = (L.TapeOp * c["n_ops"])(*c["ops"])
    for o in ops1:
        o.lane = 0
    one = lib.milb200_tape_workspace_bytes(ops1, c["n_ops"], slots, c["n_slots"], L.BF16, 0, None)
    assert two == 2 * one


def test_collapsed_fusion_program_and_segment_table(model):
    """The T = 1 CT+pathology branch as a segmented program (csrc/xfusion.cu): no projected-keys ops on the image side, the
    token side in fp32, and the native validator's view of the segment table."""
    import mil_b200
    from mil_b200 import _lib as L
    lib = mil_b200.lib()
    t = model._fusion_tape_v2()
    kinds = [o[0] for o in t.ops]
    assert kinds.count(L.OP_T2I_POOL) == 3 and kinds.count(L.OP_LN_SEG) == 2 and kinds.count(L.OP_TOK_SCATTER) == 1
    assert L.OP_ATTENTION not in kinds and L.OP_ADD not in kinds
    big = {t.inputs[0], t.inputs[1], t.inputs[2]}          # pathology rows, CT tokens, position table: program dtype
    for o in t.ops:                                         # every LINEAR except fc_pathology runs on fp32 token rows
        if o[0] == L.OP_LINEAR and o[1] not in big:
            assert t.slot_f32[o[1]] and t.slot_f32[o[4]]
    rows, segs, bag_off = model.fusion_layout(160, [300, 50], 1)
    assert bag_off == [0, 462, 674]
    assert segs[0] == ((350, 160, 1, 0), (510, 160, 463, 462), (0, 300, 162, 161), (300, 50, 624, 623))
    rows["NPE"] = 4096
    c = t._freeze()
    slots = t._slots(rows)
    segp = t._segments(segs)
    assert lib.milb200_tape_arena_bytes(c["ops"], c["n_ops"], slots, c["n_slots"], L.BF16, segp) > 670 * 512 * 2
    n = c["n_slots"]
    ext = (C.c_void_p * n)()
    for i in range(n):
        ext[i] = 4096
    dummy = C.c_void_p(4096)
    # without a segment table the segment ops are rejected before anything is launched
    l0 = lib.milb200_launch_count()
    rc = lib.milb200_tape_forward(c["ops"], c["n_ops"], slots, n, c["params"], c["n_params"], ext, dummy, dummy, dummy, 1 << 40,
                                  dummy, 1 << 40, L.BF16, None, None)
    assert rc != 0 and b"segment table" in lib.milb200_last_error()
    # a segment table whose row count contradicts the slots
    bad_segs = t._segments((segs[0][:3], 1))
    rc = lib.milb200_tape_forward(c["ops"], c["n_ops"], slots, n, c["params"], c["n_params"], ext, dummy, dummy, dummy, 1 << 40,
                                  dummy, 1 << 40, L.BF16, bad_segs, None)
    assert rc != 0 and b"mismatch" in lib.milb200_last_error()
    assert lib.milb200_launch_count() == l0


def test_c_abi_argument_checks_launch_nothing():
    """Bad arguments return an error code and a message; nothing is launched (launch counter unchanged)."""
    import mil_b200
    lib = mil_b200.lib()
    l0 = lib.milb200_launch_count()
    assert lib.milb200_segment_softmax_pool_fwd(None, None, None, 0, 0, 0, 0, None, None, None, None, None, 0, None) != 0
    assert lib.milb200_gated_score_fwd(None, None, None, None, None, None, None, 0, 0, 0, 0, None, 0, None) != 0
    assert lib.milb200_linear_fwd(None, None, None, None, None, 0, 0, 0, 0, 0, None, 0, None) != 0
    assert lib.milb200_attention_fwd(None, None, None, None, None, 0, 0, 0, 0, 0, None, 0, None) != 0
    assert lib.milb200_layernorm_fwd(None, None, None, None, None, None, None, 0, 0, 0, 0, None) != 0
    assert lib.milb200_launch_count() == l0
    assert len(lib.milb200_last_error()) > 0
