"""Extracts ``tests/golden/reference_graph.json`` from the ONE artefact the reference itself produced:
the TensorBoard GraphDef dump ``/root/reference/libs/uresnet_graph/events.out.tfevents.1509651705.lee``
(written by ``python lib/uresnet.py save``-style code of an OLDER revision of the model: 10,303 nodes,
producer 24; SURVEY.md section 3.5 lists what that revision shares with the current code and what it
does not).

The dump holds no tensors, but it does hold facts that TensorFlow-1.x wrote about the reference graph and
that a restatement can be held to:

* every filter variable's name, shape and the limits of its Xavier-uniform initializer (float32
  constants computed by ``tf.contrib.layers.xavier_initializer`` itself: known-answer values);
* every conv / conv-transpose op's strides, padding, data format and which tensor feeds it (wiring);
* how slim lowers ``conv2d_transpose`` (``Conv2DBackpropInput`` of a SAME stride-2 conv, output shape
  = 2 x input, filter ``[k,k,Cout,Cin]``);
* the concat operand order and axis, the residual adds and the ReLU behind them;
* the expansion of ``batch_norm`` (no gamma, moments over [0,1,2], biased variance, batch moments feed
  the normalisation in training, ``beta - mean * rsqrt`` form, moving-average decay);
* the accuracy wiring (argmax axis, where / gather_nd on the thresholded image).

Run in the build container (the reference tree is not on the GPU box):

    python tests/golden/make_reference_graph_fixture.py [events-file] [out.json]

``tests/test_reference_graph.py`` holds the oracle (and the library's plan) to the committed JSON and,
where ``/root/reference`` exists, re-derives the JSON from the dump and compares.
"""
import json
import os
import re
import sys

EVENTS = "/root/reference/libs/uresnet_graph/events.out.tfevents.1509651705.lee"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_graph.json")

ANCHOR_OPS = ("Conv2D", "Conv2DBackpropInput", "ConcatV2", "Placeholder")
MODULE_RELU = re.compile(r"^UResNet/(resnet_module\d+/module\d)/Relu$")


def load_graph(path):
    from tensorboard.backend.event_processing import event_file_loader
    from tensorboard.compat.proto import graph_pb2
    g = None
    for ev in event_file_loader.EventFileLoader(path).Load():
        if ev.graph_def:
            g = graph_pb2.GraphDef()
            g.ParseFromString(ev.graph_def)
    if g is None:
        raise SystemExit("no graph_def event in %s" % path)
    return g


def extract(path=EVENTS):
    from tensorboard.util import tensor_util
    g = load_graph(path)
    N = {n.name: n for n in g.node}

    def const(name):
        return tensor_util.make_ndarray(N[name].attr["value"].tensor)

    def forward(n):
        return n.name.startswith("UResNet/") and "RMSProp" not in n.name

    def anchor_name(n):
        """the name a restatement knows this tensor by"""
        m = MODULE_RELU.match(n.name)
        if m:
            return m.group(1)
        if n.op in ("Conv2D", "Conv2DBackpropInput"):
            return n.name.split("/", 1)[1].rsplit("/", 1)[0]      # UResNet/<scope>/convolution -> <scope>
        if n.op == "ConcatV2":
            return n.name.split("/", 1)[1]
        if n.op == "Placeholder":
            return {"input_prep/input_image": "data"}.get(n.name, n.name)
        raise KeyError(n.name)

    def is_anchor(n):
        return n.op in ANCHOR_OPS or bool(MODULE_RELU.match(n.name))

    def sources(name, skip_variables=True):
        """nearest anchors upstream of tensor ``name`` (in first-reached order), looking through the
        elementwise / batch-norm / reshape nodes in between"""
        out, seen, todo = [], set(), [name]
        while todo:
            cur = todo.pop(0).split(":")[0].lstrip("^")
            if cur in seen:
                continue
            seen.add(cur)
            n = N[cur]
            if n.op in ("VariableV2", "Const") or (skip_variables and n.op == "Identity" and n.input
                                                   and N[n.input[0]].op == "VariableV2"):
                continue
            if is_anchor(n):
                a = anchor_name(n)
                if a not in out:
                    out.append(a)
                continue
            todo.extend(n.input)
        return out

    fx = {"source": os.path.relpath(path, "/root/reference"), "num_nodes": len(g.node),
          "producer": int(g.versions.producer)}

    # ---- filter variables: shape + Xavier-uniform limits ---------------------------------------------
    W = {}
    for n in g.node:
        if n.op == "VariableV2" and n.name.endswith("/weights"):
            lo = float(const(n.name + "/Initializer/random_uniform/min"))
            hi = float(const(n.name + "/Initializer/random_uniform/max"))
            assert lo == -hi
            W[n.name] = {"shape": [int(d.size) for d in n.attr["shape"].shape.dim], "uniform_limit": hi}
    fx["weights"] = W

    # ---- conv-type ops ---------------------------------------------------------------------------------
    C = {}
    for n in g.node:
        if n.op in ("Conv2D", "Conv2DBackpropInput") and forward(n):
            scope = anchor_name(n)
            rec = {"op": n.op, "strides": [int(i) for i in n.attr["strides"].list.i],
                   "padding": n.attr["padding"].s.decode(), "data_format": n.attr["data_format"].s.decode()}
            if n.op == "Conv2D":
                x, w = n.input[0], n.input[1]
            else:                                   # (output_shape, filter, out_backprop = the layer's input)
                shp, w, x = n.input
                pack = N[shp]
                assert pack.op == "Pack" and len(pack.input) == 4
                mul = []
                for i in (1, 2):
                    m = N[pack.input[i]]
                    assert m.op == "Mul"
                    mul.append(int(const(m.input[1])))
                rec["output_shape"] = {"spatial_multiplier": mul, "channels": int(const(pack.input[3])),
                                       "batch_from_input": N[pack.input[0]].op == "StridedSlice"}
            assert w == "UResNet/%s/weights/read" % scope, (w, scope)
            rec["source"] = sources(x)
            C[scope] = rec
    fx["convs"] = C

    # ---- concats, residual adds ------------------------------------------------------------------------
    fx["concats"] = {}
    for n in g.node:
        if n.op == "ConcatV2" and n.name.startswith("UResNet/") and forward(n):
            fx["concats"][anchor_name(n)] = {"inputs": [sources(i)[0] for i in n.input[:-1]],
                                             "axis": int(const(n.input[-1]))}
    fx["modules"] = {}
    for n in g.node:
        m = MODULE_RELU.match(n.name)
        if m:
            add = N[n.input[0]]
            assert add.op == "Add" and len(add.input) == 2
            fx["modules"][m.group(1)] = {"relu_of": "add", "add_inputs": [sources(i)[0] for i in add.input]}

    # ---- batch_norm expansion (taken from one scope; every scope is checked to have the same form) -----
    def bn_form(pre):
        def nm(s):
            return s[len(pre):] if s.startswith(pre) else "<x>"
        form = {}
        for n in g.node:
            if n.name.startswith(pre) and "Initializer" not in n.name and n.op not in ("VariableV2", "Assign", "Identity"):
                if n.op == "Const":
                    v = const(n.name)
                    form[nm(n.name)] = ["Const", [float(x) for x in v.reshape(-1)]]
                else:
                    form[nm(n.name)] = [n.op, [nm(i) for i in n.input]]
        return form

    bn_scopes = sorted(set(n.name.rsplit("/", 1)[0] + "/" for n in g.node
                           if n.op == "VariableV2" and n.name.endswith("/beta")))
    forms = [bn_form(p) for p in bn_scopes]
    assert all(f == forms[0] for f in forms), "batch_norm scopes differ in form"
    variables = sorted(n.name[len(bn_scopes[0]):] for n in g.node
                       if n.op == "VariableV2" and n.name.startswith(bn_scopes[0]))
    fx["batch_norm"] = {"num_scopes": len(bn_scopes), "variables": variables, "form": forms[0],
                        "moving_average_consumers": sorted(
                            c.name for c in g.node
                            if any(i.lstrip("^").endswith("AssignMovingAvg") for i in c.input))}

    # ---- accuracy wiring ---------------------------------------------------------------------------------
    M = {}
    for n in g.node:
        if n.name.startswith("metrics/") and n.name.count("/") <= 2 and n.op in (
                "ArgMax", "Equal", "Cast", "Mean", "Greater", "Where", "GatherNd", "Const"):
            if n.op == "Const":
                v = const(n.name)
                if v.size <= 4:
                    M[n.name] = ["Const", [float(x) for x in v.reshape(-1)]]
            else:
                M[n.name] = [n.op, list(n.input)]
    fx["metrics"] = M
    fx["placeholder_reshapes"] = {k: [int(x) for x in const(k + "/shape")] for k in (
        "input_prep/image_reshape", "input_prep/label_reshape", "input_prep/weight_reshape")}
    return fx


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else EVENTS
    dst = sys.argv[2] if len(sys.argv) > 2 else OUT
    fx = extract(src)
    with open(dst, "w") as f:
        json.dump(fx, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote %s: %d filters, %d conv ops, %d concats, %d modules" % (
        dst, len(fx["weights"]), len(fx["convs"]), len(fx["concats"]), len(fx["modules"])))
