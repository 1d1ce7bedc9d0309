#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the scheduler hot path into a JSON fixture.

Runs in the BUILD container only (needs /root/reference); the GPU box uses the committed
``go_filter_test_vectors.json``.  Nothing is copied from the reference except the test DATA
(pod metrics literals, expected survivor lists, request fields), which is exactly what a golden
vector is.  Sources parsed:

  pkg/ext-proc/scheduling/filter_test.go   TestFilter (:12-215) and TestFilterFunc (:217-409)
  pkg/ext-proc/test/hermetic_test.go       TestHandleRequestBody (:27-139)

The parser below understands the subset of Go composite-literal syntax those tables use.

    python tests/golden/extract_go_vectors.py [--reference /root/reference] [--check]
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<str>"(?:\\.|[^"\\])*")
  | (?P<num>\d+\.\d*|\.\d+|\d+)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>:=|!=|==|&&|\|\||[{}\[\](),:*.&=<>!+\-/;%|])
""", re.X)


def tokenize(src: str):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at line {line}: {src[pos:pos+30]!r}")
        kind = m.lastgroup
        text = m.group()
        if kind != "ws":
            toks.append((kind, text, line))
        line += text.count("\n")
        pos = m.end()
    return toks


class Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "", -1)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, text):
        tok = self.next()
        if tok[1] != text:
            raise SyntaxError(f"line {tok[2]}: expected {text!r}, got {tok[1]!r}")
        return tok

    def skip_balanced(self, open_, close):
        depth = 0
        while True:
            tok = self.next()
            if tok[0] == "eof":
                raise SyntaxError("unbalanced")
            if tok[1] == open_:
                depth += 1
            elif tok[1] == close:
                depth -= 1
                if depth == 0:
                    return

    def parse_value(self):
        kind, text, line = self.peek()
        if kind == "str":
            self.next()
            return json.loads(text)
        if kind == "num":
            self.next()
            return float(text) if "." in text else int(text)
        if text == "&":
            self.next()
            return self.parse_value()
        if text == "func":
            # func(params) (results) { body }  -> opaque
            start_line = line
            self.next()
            self.skip_balanced("(", ")")
            if self.peek()[1] == "(":
                self.skip_balanced("(", ")")
            else:
                while self.peek()[1] != "{":
                    self.next()
            self.skip_balanced("{", "}")
            return {"$func": f"line {start_line}"}
        if text == "{":
            return self.parse_composite(None, line)
        # a type / identifier / call
        parts = []
        while True:
            kind, text, _ = self.peek()
            if kind == "id" or text in (".", "[", "]", "*"):
                parts.append(text)
                self.next()
            else:
                break
        name = "".join(parts)
        nxt = self.peek()[1]
        if nxt == "{":
            return self.parse_composite(name, line)
        if nxt == "(":
            self.next()
            args = []
            while self.peek()[1] != ")":
                args.append(self.parse_value())
                if self.peek()[1] == ",":
                    self.next()
            self.expect(")")
            return {"$call": name, "args": args, "line": line}
        if name == "true":
            return True
        if name == "false":
            return False
        if name == "nil":
            return None
        return {"$ident": name}

    def parse_composite(self, type_name, line):
        self.expect("{")
        keyed, items = {}, []
        while self.peek()[1] != "}":
            # keyed element?  (identifier or string followed by ':')
            k0, k1 = self.peek(), self.peek(1)
            if k1[1] == ":" and k0[0] in ("id", "str"):
                key = json.loads(k0[1]) if k0[0] == "str" else k0[1]
                self.next()
                self.next()
                keyed[key] = self.parse_value()
            else:
                items.append(self.parse_value())
            if self.peek()[1] == ",":
                self.next()
        self.expect("}")
        out = {"$type": type_name, "$line": line}
        if keyed:
            out["fields"] = keyed
        if items or not keyed:
            out["items"] = items
        return out


def find_table(toks, func_name):
    """Return the parsed `tests := []struct{...}{ ... }` literal inside `func func_name`."""
    i = 0
    while i < len(toks):
        if toks[i][1] == "func" and toks[i + 1][1] == func_name:
            break
        i += 1
    else:
        raise SyntaxError(f"func {func_name} not found")
    while not (toks[i][1] == "tests" and toks[i + 1][1] == ":="):
        i += 1
    i += 2
    # []struct { field decls } { elements }
    assert toks[i][1] == "[" and toks[i + 2][1] == "struct", toks[i:i + 4]
    p = Parser(toks)
    p.i = i + 3
    p.skip_balanced("{", "}")
    return p.parse_composite("tests", toks[i][2])


def pod_metrics(node, fake_pod=False):
    """backend.PodMetrics literal -> plain dict (zero values filled in like Go does)."""
    f = node.get("fields", {})
    pod = {"name": "", "address": ""}
    if "Pod" in f:
        pv = f["Pod"]
        if "$call" in pv and pv["$call"] == "FakePod":
            # test/utils.go:73-80: Name "pod-<i>", Address "address-<i>"
            idx = pv["args"][0]
            pod = {"name": f"pod-{idx}", "address": f"address-{idx}"}
        else:
            pf = pv.get("fields", {})
            pod = {"name": pf.get("Name", ""), "address": pf.get("Address", "")}
    m = f.get("Metrics", {"fields": {}}).get("fields", {})
    active = m.get("ActiveModels", {"fields": {}}).get("fields", {})
    return {
        "name": pod["name"],
        "address": pod["address"],
        "waiting_queue_size": m.get("WaitingQueueSize", 0),
        "kv_cache_usage_percent": float(m.get("KVCacheUsagePercent", 0)),
        "max_active_models": m.get("MaxActiveModels", 0),
        "active_models": sorted(active.keys()),
        "line": node["$line"],
    }


def pod_list(node):
    if node is None:
        return None
    return [pod_metrics(x) for x in node.get("items", [])]


def request(node):
    if node is None:
        return None
    f = node.get("fields", {})
    return {
        "model": f.get("Model", ""),
        "resolved_target_model": f.get("ResolvedTargetModel", ""),
        "critical": bool(f.get("Critical", False)),
    }


def describe_func(v):
    if v is None:
        return None
    if "$ident" in v:
        return {"name": v["$ident"]}
    if "$call" in v:
        return {"name": v["$call"], "args": [describe_func(a) if isinstance(a, dict) else a
                                             for a in v["args"]]}
    if "$func" in v:
        return {"name": "$func_literal"}
    if "$type" in v:  # &filter{filter: func...}
        inner = v.get("fields", {})
        return {"name": "$filter_literal",
                "fields": {k: describe_func(x) if isinstance(x, dict) else x
                           for k, x in inner.items()}}
    return v


def extract(reference_root: str):
    base = os.path.join(reference_root, "pkg", "ext-proc")
    rel_ft = "pkg/ext-proc/scheduling/filter_test.go"
    toks = tokenize(open(os.path.join(base, "scheduling", "filter_test.go")).read())
    out = {"source_commit": "8e96339eedd63991ebf1622f326cc25f9115a6aa", "TestFilter": [],
           "TestFilterFunc": [], "TestHandleRequestBody": []}
    for case in find_table(toks, "TestFilter")["items"]:
        f = case["fields"]
        out["TestFilter"].append({
            "name": f["name"], "source": f"{rel_ft}:{case['$line']}",
            "filter": describe_func(f.get("filter")),
            "req": request(f.get("req")),
            "input": pod_list(f.get("input")) or [],
            "output": pod_list(f.get("output")),       # None = Go nil slice
            "err": bool(f.get("err", False)),
        })
    for case in find_table(toks, "TestFilterFunc")["items"]:
        f = case["fields"]
        out["TestFilterFunc"].append({
            "name": f["name"], "source": f"{rel_ft}:{case['$line']}",
            "f": describe_func(f.get("f")),
            "req": request(f.get("req")),
            "input": pod_list(f.get("input")) or [],
            "output": pod_list(f.get("output")),
            "err": bool(f.get("err", False)),
        })
    rel_ht = "pkg/ext-proc/test/hermetic_test.go"
    toks = tokenize(open(os.path.join(base, "test", "hermetic_test.go")).read())
    for case in find_table(toks, "TestHandleRequestBody")["items"]:
        f = case["fields"]
        models = {}
        for mname, mv in f["models"]["fields"].items():
            spec = mv["fields"]["Spec"]["fields"]
            targets = [{"name": t["fields"]["Name"], "weight": t["fields"].get("Weight", 0)}
                       for t in spec.get("TargetModels", {"items": []})["items"]]
            models[mname] = {"model_name": spec["ModelName"], "target_models": targets,
                             "criticality": None if "Criticality" not in spec else
                             describe_func(spec["Criticality"])}
        headers = []
        for h in f["wantHeaders"]["items"]:
            hv = h["fields"]["Header"]["fields"]
            headers.append({"key": hv["Key"], "raw_value": hv["RawValue"]["args"][0]})
        out["TestHandleRequestBody"].append({
            "name": f["name"], "source": f"{rel_ht}:{case['$line']}",
            "request_model": f["req"]["args"][0],
            "models": models,
            "pods": [pod_metrics(x) for x in f["pods"]["items"]],
            "want_headers": headers,
            "want_body": f["wantBody"]["args"][0],
            "want_err": bool(f.get("wantErr", False)),
        })
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--check", action="store_true",
                    help="fail if the committed fixture differs from a fresh extraction")
    args = ap.parse_args()
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "go_filter_test_vectors.json")
    data = extract(args.reference)
    text = json.dumps(data, indent=1, sort_keys=True) + "\n"
    if args.check:
        if open(path).read() != text:
            print("fixture out of date", file=sys.stderr)
            return 1
        print("fixture matches the reference")
        return 0
    with open(path, "w") as fh:
        fh.write(text)
    print(f"wrote {path}: {len(data['TestFilter'])} TestFilter, "
          f"{len(data['TestFilterFunc'])} TestFilterFunc, "
          f"{len(data['TestHandleRequestBody'])} hermetic cases")
    return 0


if __name__ == "__main__":
    sys.exit(main())
