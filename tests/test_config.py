"""Host-side compile step (bqa_b200/config.py) against the reference's golden layout vectors
(reference tests/test_config_to_context.py:36-164) and against the oracle's loop-form compiler."""
from math import isclose

import numpy as np
import pytest

import instances
from bqa_b200.config import ConfigSyntaxError, config_to_context
from oracle import bqa_oracle as O

REF_TEST_CONFIG = {   # reference tests/test_config_to_context.py:38-60
    "edges": {(0, 2): 1.0, (1, 0): -1, (3, 0): 0.1, (2, 4): 1.1, (1, 4): 0, (3, 1): 1},
    "nodes": {2: 1, 6: -1.1},
    "default_field": -0.5,
    "schedule": {
        "starting_mixing": 0.8, "total_time": 5,
        "actions": [
            {"weight": 0.4, "final_mixing": 0.3, "steps_number": 8},
            "measure",
            {"type": "imag_time_evolution", "weight": 0.6, "final_mixing": 0.11, "steps_number": 10},
            "get_bloch_vectors",
        ],
    },
    "damping": 0.3,
}


def test_reference_golden_layouts():
    ctx = config_to_context(REF_TEST_CONFIG)
    assert ctx.edges_number == 12 and ctx.nodes_number == 7
    assert ctx.max_bp_iters_number == 75 and ctx.max_bond_dim == 4
    assert isclose(ctx.bp_eps, 1e-6) and isclose(ctx.damping, 0.3)
    assert list(ctx.degree_to_layout) == [3, 2, 0]          # first-appearance order by node id
    l0, l2, l3 = ctx.degree_to_layout[0], ctx.degree_to_layout[2], ctx.degree_to_layout[3]
    # reference tests/test_config_to_context.py:68-75
    assert l0.node_ids.tolist() == [5, 6] and np.allclose(l0.node_ampls, [-0.5, -1.1])
    assert l0.input_msgs_position.shape == (0, 2)
    # :76-97
    assert l2.node_ids.tolist() == [2, 3, 4]
    assert l2.input_msgs_position.tolist() == [[9, 8, 3], [0, 11, 4]]
    assert l2.output_msgs_position.tolist() == [[3, 2, 9], [6, 5, 10]]
    assert l2.lmbds_position.tolist() == [[3, 2, 3], [0, 5, 4]]
    assert np.allclose(l2.node_ampls, [1., -0.5, -0.5])
    assert np.allclose(l2.edge_ampls, [[1.1, 0.1, 1.1], [1., 1., 0.]])
    # :98-119
    assert l3.node_ids.tolist() == [0, 1]
    assert l3.input_msgs_position.tolist() == [[6, 7], [1, 10], [2, 5]]
    assert l3.output_msgs_position.tolist() == [[0, 1], [7, 4], [8, 11]]
    assert l3.lmbds_position.tolist() == [[0, 1], [1, 4], [2, 5]]
    assert np.allclose(l3.node_ampls, [-0.5, -0.5])
    assert np.allclose(l3.edge_ampls, [[1., -1.], [-1., 0.], [0.1, 1.]])
    for node_id, (degree, pos) in ctx.path_to_tensors.items():
        assert ctx.degree_to_layout[degree].node_ids[pos] == node_id


def test_reference_golden_schedule():
    # reference tests/test_config_to_context.py:120-162
    ins = config_to_context(REF_TEST_CONFIG).instructions
    assert len(ins) == 8 + 1 + 10 + 1 and ins[8] == "measure" and ins[-1] == "get_bloch_vectors"
    layers = [i for i in ins if isinstance(i, dict)]
    assert isclose(sum(i["xtime"] + i["ztime"] for i in layers), 5.0)
    assert isclose(sum(i["xtime"] + i["ztime"] for i in layers[:8]), 2.0)
    mix = [i["xtime"] / (i["xtime"] + i["ztime"]) for i in layers]
    assert isclose(mix[0], 0.8) and isclose(mix[8], 0.3)
    d1 = np.diff(mix[:8] + [mix[8]])
    d2 = np.diff(mix[8:] + [0.11])
    assert np.allclose(d1, d1[0]) and np.allclose(d2, d2[0])
    assert layers[8]["type"] == "imag_time_evolution"


@pytest.mark.parametrize("name", list(instances.GOLDEN_CONFIGS))
def test_layouts_equal_oracle_compiler(name):
    cfg = instances.GOLDEN_CONFIGS[name]()
    ctx = config_to_context(cfg)
    octx = O.compile_config(cfg)
    assert list(ctx.degree_to_layout) == list(octx.layouts)
    for d, lay in ctx.degree_to_layout.items():
        ol = octx.layouts[d]
        assert lay.node_ids.tolist() == ol.node_ids.tolist()
        for j in range(d):
            assert lay.input_msgs_position[j].tolist() == ol.in_pos[j].tolist()
            assert lay.output_msgs_position[j].tolist() == ol.out_pos[j].tolist()
            assert lay.lmbds_position[j].tolist() == ol.lmbd_pos[j].tolist()
            assert np.allclose(lay.edge_ampls[j], ol.edge_ampls[j].real)
        assert np.allclose(lay.node_ampls, ol.node_ampls.real)
    assert ctx.graph == octx.graph
    for a, b in zip(ctx.instructions, octx.instructions):
        assert a == b or (isclose(a["xtime"], b["xtime"]) and isclose(a["ztime"], b["ztime"]))


def test_edges_as_list_and_defaults():
    ctx = config_to_context({"edges": [((0, 1), 0.5), [(1, 2), -1]]})
    assert ctx.nodes_number == 3 and ctx.edges_number == 4
    assert len(ctx.instructions) == 101 and ctx.instructions[-1] == "get_bloch_vectors"
    assert isclose(ctx.instructions[0]["ztime"], 0.0) and isclose(ctx.instructions[0]["xtime"], 0.1)


@pytest.mark.parametrize("bad", [
    "not a dict",
    {},                                                            # edges missing
    {"edges": {(0, 0): 1.0}},                                      # self loop
    {"edges": {(0, 1): 1.0, (1, 0): 2.0}},                         # duplicated edge
    {"edges": {(0, 1): "x"}},
    {"edges": {(0, 1): 1.0}, "max_bond_dim": 0},
    {"edges": {(0, 1): 1.0}, "measurement_threshold": 0.3},
    {"edges": {(0, 1): 1.0}, "damping": 1.5},
    {"edges": {(0, 1): 1.0}, "backend": "tpu"},
    {"edges": {(0, 1): 1.0}, "nodes": {-1: 0.5}},
    {"edges": {(0, 1): 1.0}, "schedule": {"actions": [{"weight": 0.5}]}},           # weights do not sum to 1
    {"edges": {(0, 1): 1.0}, "schedule": {"actions": [{"weight": 1.0, "initial_mixing": 0.2}]}},
    {"edges": {(0, 1): 1.0}, "schedule": {"actions": [{"weight": 1.0}, "explode"]}},
])
def test_syntax_errors(bad):
    with pytest.raises(ConfigSyntaxError):
        config_to_context(bad)


def test_unknown_keys_ignored():
    # e.g. the misspelt `max_bp_iters_number` is dropped silently
    ctx = config_to_context({"edges": {(0, 1): 1.0}, "max_bp_iters_number": 7, "runtime_limit": 3})
    assert ctx.max_bp_iters_number == 75
