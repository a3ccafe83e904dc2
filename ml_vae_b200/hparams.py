"""PyYAML-only loader for the HyperPyYAML subset the reference's experiment files use.

The reference vendors a modified hyperpyyaml (src/hyperpyyaml/core.py) that needs
ruamel.yaml; this loader covers exactly the constructs its run.yaml / model.yaml files
contain, so the same files load where ruamel is absent:

  !new:pkg.Class {kwargs} / [args]      construct an object          (core.py:428-444)
  !name:pkg.func {kwargs}               functools.partial             (core.py:447-467)
  !apply:pkg.func [args] / {kwargs}     call now, keep the result     (core.py:486-502)
  !ref <a.b> , "text <a>", <a> * 3      reference / interpolation / arithmetic (core.py:505-661)
  !copy <a>                             deep copy of the referenced node
  !include:relative/path.yaml           nested file; child keys of the node are passed
                                        down as overrides of the included file
  !PLACEHOLDER                          must be overridden
  overrides                             dict, yaml string, or a list of those, applied in order
                                        (fork feature, core.py:308-312)
  tagged scalar + child keys            `model: !PLACEHOLDER` followed by children
                                        (config/run.yaml:60-66): the override replaces the tag,
                                        the children stay and are injected (core.py:704-717)

Objects are built in document order (so `__set_seed: !apply:torch.manual_seed` at
run.yaml:3 runs before any module is constructed) and every key is constructed once:
`!ref <encoder>` yields the same instance wherever it appears.
"""
from __future__ import annotations

import ast
import copy
import functools
import operator as op
import os
import pydoc
import re

import yaml

_REF = re.compile(r"<([^<>\s]+)>")


class _Tagged:
    __slots__ = ("kind", "arg", "value")

    def __init__(self, kind, arg, value):
        self.kind, self.arg, self.value = kind, arg, value

    def __repr__(self):
        return f"_Tagged({self.kind}:{self.arg}, {self.value!r})"


def _to_tree(node):
    """yaml node graph -> plain dict / list / scalar tree with _Tagged wrappers."""
    tag = node.tag
    if isinstance(node, yaml.MappingNode):
        val = {_to_tree(k): _to_tree(v) for k, v in node.value}
    elif isinstance(node, yaml.SequenceNode):
        val = [_to_tree(v) for v in node.value]
    else:
        if tag.startswith("!"):
            val = node.value
        else:
            val = yaml.SafeLoader.construct_object(_SCALAR_LOADER, node)
    if tag.startswith("!") and not tag.startswith("!!"):
        body = tag[1:]
        kind, _, arg = body.partition(":")
        if kind == "tuple":
            return tuple(val)
        return _Tagged(kind, arg, val if val != "" else None)
    return val


class _ScalarLoader(yaml.SafeLoader):
    pass


_SCALAR_LOADER = _ScalarLoader("")


def _parse(stream):
    text = stream.read() if hasattr(stream, "read") else stream
    node = yaml.compose(text, Loader=yaml.SafeLoader)
    return {} if node is None else _to_tree(node)


def _merge(tree, upd):
    """recursive update; a tagged node that receives a dict keeps the dict as its children."""
    for k, v in upd.items():
        cur = tree.get(k)
        if isinstance(v, dict) and isinstance(cur, dict):
            _merge(cur, v)
        elif isinstance(v, dict) and isinstance(cur, _Tagged) and isinstance(cur.value, dict):
            _merge(cur.value, v)
        elif isinstance(v, _Tagged) and v.value is None and isinstance(cur, _Tagged) and isinstance(cur.value, dict):
            tree[k] = _Tagged(v.kind, v.arg, cur.value)          # replace the tag, keep injected children
        else:
            tree[k] = v
    return tree


def _as_override(o):
    if o is None or o == "":
        return {}
    if isinstance(o, str):
        return _parse(o)

    def shallow(d):      # copy the dict skeleton only: leaves may be live objects (identity matters)
        return {k: shallow(v) if isinstance(v, dict) else v for k, v in d.items()}
    return shallow(o)


_OPS = {ast.Add: op.add, ast.Sub: op.sub, ast.Mult: op.mul, ast.Div: op.truediv, ast.FloorDiv: op.floordiv,
        ast.Pow: op.pow, ast.Mod: op.mod, ast.USub: op.neg}


def _arith(node):
    if isinstance(node, ast.Constant):
        return node.value
    if isinstance(node, ast.BinOp):
        return _OPS[type(node.op)](_arith(node.left), _arith(node.right))
    if isinstance(node, ast.UnaryOp):
        return _OPS[type(node.op)](_arith(node.operand))
    raise ValueError("unsupported arithmetic in !ref")


class _Builder:
    def __init__(self, tree, base_dir):
        self.tree, self.base_dir = tree, base_dir
        self.built = {}            # key path -> constructed python object
        self.busy = set()

    # ---- references ------------------------------------------------------------------
    def _node_at(self, path):
        cur = self.tree
        for part in path.split("."):
            if isinstance(cur, _Tagged):
                cur = cur.value
            if isinstance(cur, list):
                cur = cur[int(part)]
            else:
                if part not in cur:
                    raise KeyError(f"!ref <{path}>: no such key")
                cur = cur[part]
        return cur

    def get(self, path):
        if path in self.built:
            return self.built[path]
        if path in self.busy:
            raise ValueError(f"circular reference through <{path}>")
        self.busy.add(path)
        try:
            obj = self.construct(self._node_at(path), path)
        finally:
            self.busy.discard(path)
        self.built[path] = obj
        return obj

    def _ref(self, text, deep=False):
        text = str(text).strip()
        m = _REF.fullmatch(text)
        if m:
            v = self.get(m.group(1))
            return copy.deepcopy(v) if deep else v
        parts = {p: self.get(p) for p in _REF.findall(text)}
        out = _REF.sub(lambda mm: str(parts[mm.group(1)]), text)
        if parts and all(isinstance(v, (int, float)) for v in parts.values()) and re.search(r"[-+*/%]", out):
            try:
                return _arith(ast.parse(out, mode="eval").body)
            except Exception:
                pass
        return out

    # ---- construction ----------------------------------------------------------------
    def construct(self, node, path=""):
        if isinstance(node, dict):
            out = {}
            for k, v in node.items():
                sub = f"{path}.{k}" if path else str(k)
                out[k] = self.get(sub)
            return out
        if isinstance(node, (list, tuple)):
            seq = [self.get(f"{path}.{i}" if path else str(i)) for i in range(len(node))]
            return tuple(seq) if isinstance(node, tuple) else seq
        if not isinstance(node, _Tagged):
            return node
        kind, arg, val = node.kind, node.arg, node.value
        if kind == "PLACEHOLDER":
            raise ValueError(f"'{path}' is a !PLACEHOLDER and must be overridden")
        if kind in ("ref", "copy"):
            return self._ref(val, deep=(kind == "copy"))
        if kind == "include":
            fname = os.path.join(self.base_dir, arg)
            children = val if isinstance(val, dict) else {}
            injected = {k: self.get(f"{path}.{k}") for k in children}     # built in THIS file's context
            with open(fname) as f:
                return load_hyperpyyaml(f, overrides=injected, _base_dir=os.path.dirname(os.path.abspath(fname)))
        args, kwargs = self._call_args(val, path)
        target = pydoc.locate(arg)
        if target is None:
            raise ImportError(f"{path}: cannot locate '{arg}'")
        if kind == "new" or kind == "apply":
            return target(*args, **kwargs)
        if kind == "name":
            return functools.partial(target, *args, **kwargs)
        if kind == "module":
            return target
        raise ValueError(f"{path}: unsupported tag !{kind}")

    def _call_args(self, val, path):
        if val is None:
            return [], {}
        if isinstance(val, dict):
            return [], {k: self.get(f"{path}.{k}") for k in val}
        if isinstance(val, (list, tuple)):
            return [self.get(f"{path}.{i}") for i in range(len(val))], {}
        return [val], {}


def load_hyperpyyaml(yaml_stream, overrides=None, overrides_must_match=True, _base_dir=None):
    """Same call signature as the reference's loader (src/hyperpyyaml/core.py:25):
    stream (or text) + overrides -> dict of constructed objects."""
    base = _base_dir or os.path.dirname(getattr(yaml_stream, "name", "")) or os.getcwd()
    tree = _parse(yaml_stream)
    ovs = overrides if isinstance(overrides, (list, tuple)) else [overrides]
    for o in ovs:
        _merge(tree, _as_override(o))
    return _Builder(tree, base).construct(tree)


def recursive_update(d, u, must_match=False):
    """src/hyperpyyaml/core.py:664-717 (post-load application of extra_overrides,
    prepare_experiment.py:25)."""
    for k, v in u.items():
        if isinstance(v, dict) and isinstance(d.get(k), dict):
            recursive_update(d[k], v, must_match)
        elif must_match and k not in d:
            raise KeyError(f"Override '{k}' not found in: {list(d)}")
        else:
            d[k] = v
