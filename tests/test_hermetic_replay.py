"""Replay of the reference's end-to-end test through the GPU scheduler behind the handler surface:
pkg/ext-proc/test/hermetic_test.go:27-139 (the request GenerateRequest builds, test/utils.go:53-70;
the expected headers and body :91-104), the 429 mapping of handlers/server.go:97-109 and the
stream-closing errors of :110-112.  The swap point is the one-line change of main.go:137 /
test/utils.go:45: scheduling.NewScheduler(pp) -> the GPU scheduler."""
import json

import pytest

from helpers import golden_to_podmetrics
from llm_instance_gateway_b200.backend import (CRITICAL, FakeDataStore, InferenceModel, InferenceModelSpec, Metrics, Pod,
                                               PodMetrics, TargetModel)
from llm_instance_gateway_b200.handlers import NewServer, RequestContext, go_json_marshal
from llm_instance_gateway_b200.scheduling import NewScheduler, StatusError

pytestmark = pytest.mark.gpu


class StaticProvider:
    def __init__(self, pods):
        self.pods = list(pods)

    def AllPodMetrics(self):
        return list(self.pods)


def generate_request(model):            # test/utils.go:53-70 (json.Marshal sorts the map keys)
    return {"request_body": {"body": go_json_marshal({"model": model, "prompt": "hello", "max_tokens": 100, "temperature": 0})}}


def test_go_json_marshal_matches_encoding_json():
    assert go_json_marshal({"model": "m", "prompt": "hello", "max_tokens": 100, "temperature": 0}) == \
        b'{"max_tokens":100,"model":"m","prompt":"hello","temperature":0}'
    assert go_json_marshal({"t": 0.7, "n": 1e21, "a": [1.5, True, None], "s": "<&>"}) == \
        b'{"a":[1.5,true,null],"n":1e+21,"s":"\\u003c\\u0026\\u003e","t":0.7}'


def test_hermetic_success_case(golden):
    case = golden["TestHandleRequestBody"][0]
    pods = [golden_to_podmetrics(p) for p in case["pods"]]
    models = {name: InferenceModel(name, InferenceModelSpec(
        ModelName=name, TargetModels=[TargetModel(t["name"], t["weight"]) for t in m["target_models"]]))
        for name, m in case["models"].items()}
    pp = StaticProvider(pods)
    sched = NewScheduler(pp, max_pods=64, max_adapters=64, max_batch=64)
    try:
        server = NewServer(pp, sched, "target-pod", FakeDataStore(models))
        ctx = RequestContext()
        assert server.Process(ctx, {"request_headers": {}}) == {"request_headers": {"response": {"clear_route_cache": True}}}
        res = server.Process(ctx, generate_request(case["request_model"]))
        want_headers = [{"header": {"key": h["key"], "raw_value": h["raw_value"].encode()}} for h in case["want_headers"]]
        assert res == {"request_body": {"response": {"header_mutation": {"set_headers": want_headers},
                                                     "body_mutation": {"body": case["want_body"].encode()}}}}
        assert want_headers[0]["header"] == {"key": "target-pod", "raw_value": b"address-1"}
        assert want_headers[1]["header"] == {"key": "Content-Length", "raw_value": b"73"}
        assert ctx.TargetPod == Pod("pod-1", "address-1") and ctx.Model == "my-model"
    finally:
        sched.close()


def test_shed_is_429_and_other_errors_close_the_stream():
    busy = [PodMetrics(Pod(f"pod-{i}", f"address-{i}"), Metrics(WaitingQueueSize=10, KVCacheUsagePercent=0.9)) for i in range(3)]
    models = {"sheddable": InferenceModel("sheddable", InferenceModelSpec(ModelName="sheddable")),
              "critical": InferenceModel("critical", InferenceModelSpec(ModelName="critical", Criticality=CRITICAL)),
              "split": InferenceModel("split", InferenceModelSpec(ModelName="split", Criticality=CRITICAL,
                                                                  TargetModels=[TargetModel("canary", 50), TargetModel("v1", 50)]))}
    pp = StaticProvider(busy)
    sched = NewScheduler(pp, max_pods=64, max_adapters=64, max_batch=64)
    try:
        server = NewServer(pp, sched, "target-pod", FakeDataStore(models), seed=11)
        # no capacity for a sheddable request: ResourceExhausted -> ImmediateResponse 429   server.go:97-109
        assert server.Process(RequestContext(), generate_request("sheddable")) == {"immediate_response": {"status": {"code": 429}}}
        # the same pool still serves a critical one, body untouched when the model name is kept
        res = server.Process(RequestContext(), generate_request("critical"))
        body = res["request_body"]["response"]["body_mutation"]["body"]
        assert json.loads(body)["model"] == "critical"
        assert res["request_body"]["response"]["header_mutation"]["set_headers"][1]["header"]["raw_value"] == str(len(body)).encode()
        # weighted split: both targets get drawn, the body carries the drawn name
        seen = set()
        for _ in range(40):
            res = server.Process(RequestContext(), generate_request("split"))
            seen.add(json.loads(res["request_body"]["response"]["body_mutation"]["body"])["model"])
        assert seen == {"canary", "v1"}
        # unknown model / bad body: the stream is closed with the error   server.go:110-112
        with pytest.raises(StatusError) as ei:
            server.Process(RequestContext(), generate_request("nope"))
        assert ei.value.code == "Unknown" and "error finding a model object in InferenceModel for input nope" in str(ei.value)
        with pytest.raises(StatusError) as ei:
            server.Process(RequestContext(), {"request_body": {"body": b"{"}})
        assert "error unmarshaling request body" in str(ei.value)
        with pytest.raises(StatusError) as ei:
            server.Process(RequestContext(), {"request_body": {"body": b'{"prompt": "x"}'}})
        assert "model not found in request" in str(ei.value)
    finally:
        sched.close()
