"""REST text-generation server (parity: megatron/text_generation_server.py:17-241).

``PUT /api`` with the same JSON fields as the reference (prompts <= 128, tokens_to_generate, logprobs, temperature,
top_k, top_p, top_p_decay, top_p_bound, add_BOS, stop_on_double_eol, stop_on_eol, prevent_newline_after_colon,
random_seed, no_log, beam_width, stop_token, length_penalty).  Rank 0 serves HTTP and broadcasts an op code so every
rank enters ``generate`` / ``beam_search``.  Flask is not a dependency: the server is a schema-validated handler on
the standard library's ``ThreadingHTTPServer`` (requests are serialised by a lock, as in the reference)."""
from __future__ import annotations

import datetime
import json
import os
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import torch
import torch.distributed as dist

from .text_generation import beam_search_and_post_process, generate_and_post_process
from .utils.device import current_device

GENERATE_NUM = 0
BEAM_NUM = 1
lock = threading.Lock()


def send_do_generate():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(torch.tensor([GENERATE_NUM], dtype=torch.long, device=current_device()), 0)


def send_do_beam_search():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(torch.tensor([BEAM_NUM], dtype=torch.long, device=current_device()), 0)


class RequestError(Exception):
    def __init__(self, message, status=400):
        super().__init__(message)
        self.message, self.status = message, status


def _number(v):
    return type(v) in (int, float)


def parse_request(req: dict) -> dict:
    """Validate a request body; returns the normalised parameter dict or raises RequestError."""
    if "prompts" not in req:
        raise RequestError("prompts argument required")
    if "max_len" in req:
        raise RequestError("max_len is no longer used.  Replace with tokens_to_generate")
    if "sentences" in req:
        raise RequestError("sentences is no longer used.  Replace with prompts")
    prompts = req["prompts"]
    if not isinstance(prompts, list):
        raise RequestError("prompts is not a list of strings")
    if len(prompts) == 0:
        raise RequestError("prompts is empty")
    if len(prompts) > 128:
        raise RequestError("Maximum number of prompts is 128")
    p = dict(prompts=prompts, tokens_to_generate=64, logprobs=False, temperature=1.0, top_k=0, top_p=0.0,
             top_p_decay=0.0, top_p_bound=0.0, add_BOS=False, stop_on_double_eol=False, stop_on_eol=False,
             prevent_newline_after_colon=False, random_seed=-1, no_log=False, beam_width=None, stop_token=50256,
             length_penalty=1)

    def take(name, check, message):
        if name in req:
            if not check(req[name]):
                raise RequestError(message)
            p[name] = req[name]

    take("tokens_to_generate", lambda v: isinstance(v, int) and not isinstance(v, bool) and v >= 0,
         "tokens_to_generate must be an integer greater than or equal to 0")
    take("logprobs", lambda v: isinstance(v, bool), "logprobs must be a boolean value")
    if p["tokens_to_generate"] == 0 and not p["logprobs"]:
        raise RequestError("tokens_to_generate=0 implies logprobs should be True")
    take("temperature", lambda v: _number(v) and 0.0 < v <= 100.0,
         "temperature must be a positive number less than or equal to 100.0")
    take("top_k", lambda v: type(v) == int and 0 <= v <= 1000,
         "top_k must be an integer equal to or greater than 0 and less than or equal to 1000")
    take("top_p", lambda v: type(v) == float and 0 <= v <= 1.0, "top_p must be a positive float less than or equal to 1.0")
    if p["top_p"] > 0.0 and p["top_k"] > 0:
        raise RequestError("cannot set both top-k and top-p samplings.")
    take("top_p_decay", lambda v: type(v) == float and 0 <= v <= 1.0,
         "top_p_decay must be a positive float less than or equal to 1.0")
    if "top_p_decay" in req and p["top_p"] == 0.0:
        raise RequestError("top_p_decay cannot be set without top_p")
    take("top_p_bound", lambda v: type(v) == float and 0.0 < v <= p["top_p"],
         "top_p_bound must be greater than 0 and less than top_p")
    if "top_p_bound" in req and p["top_p"] == 0.0:
        raise RequestError("top_p_bound cannot be set without top_p")
    for name in ("add_BOS", "stop_on_double_eol", "stop_on_eol", "prevent_newline_after_colon", "no_log"):
        take(name, lambda v: isinstance(v, bool), f"{name} must be a boolean value")
    if any(len(x) == 0 for x in prompts) and not p["add_BOS"]:
        raise RequestError("Empty prompts require add_BOS=true")
    take("random_seed", lambda v: isinstance(v, int) and v >= 0,
         "random_seed must be a positive integer")
    take("beam_width", lambda v: isinstance(v, int) and v >= 1, "beam_width must be an integer > 1")
    if p["beam_width"] is not None and len(prompts) > 1:
        raise RequestError("When doing beam_search, batch size must be 1")
    take("stop_token", lambda v: isinstance(v, int), "stop_token must be an integer")
    take("length_penalty", _number, "length_penalty must be a float")
    return p


class MegatronGenerate:
    """The ``/api`` resource: ``put(request_dict)`` -> (json-able body, status)."""

    def __init__(self, model):
        self.model = model

    send_do_generate = staticmethod(send_do_generate)          # (static methods in the reference, :21-29)
    send_do_beam_search = staticmethod(send_do_beam_search)

    def put(self, req: dict):
        try:
            p = parse_request(req)
        except RequestError as e:
            return e.message, e.status
        with lock:
            if not p["no_log"]:
                print("request IP: local", flush=True)
                print(json.dumps(req), flush=True)
                print("start time: ", datetime.datetime.now(), flush=True)
            try:
                if p["beam_width"] is not None:
                    send_do_beam_search()
                    response, response_seg, response_scores = beam_search_and_post_process(
                        self.model, prompts=p["prompts"], tokens_to_generate=p["tokens_to_generate"],
                        beam_size=p["beam_width"], add_BOS=p["add_BOS"], stop_token=p["stop_token"],
                        num_return_gen=p["beam_width"], length_penalty=p["length_penalty"],
                        prevent_newline_after_colon=p["prevent_newline_after_colon"])
                    return {"text": response, "segments": response_seg, "scores": response_scores}, 200
                send_do_generate()
                response, response_seg, response_logprobs, _ = generate_and_post_process(
                    self.model, prompts=p["prompts"], tokens_to_generate=p["tokens_to_generate"],
                    return_output_log_probs=p["logprobs"], top_k_sampling=p["top_k"], top_p_sampling=p["top_p"],
                    top_p_decay=p["top_p_decay"], top_p_bound=p["top_p_bound"], temperature=p["temperature"],
                    add_BOS=p["add_BOS"], use_eod_token_for_early_termination=True,
                    stop_on_double_eol=p["stop_on_double_eol"], stop_on_eol=p["stop_on_eol"],
                    prevent_newline_after_colon=p["prevent_newline_after_colon"], random_seed=p["random_seed"])
                return {"text": response, "segments": response_seg, "logprobs": response_logprobs}, 200
            except ValueError as ve:
                return ve.args[0], 400


_INDEX_HTML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "static", "index.html")


class MegatronServer:
    def __init__(self, model):
        self.resource = MegatronGenerate(model)
        resource = self.resource

        class Handler(BaseHTTPRequestHandler):
            def _send(self, body, status=200, ctype="application/json"):
                data = body if isinstance(body, bytes) else (json.dumps(body) if not isinstance(body, str) else body).encode()
                self.send_response(status)
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            def do_GET(self):
                if self.path in ("/", "/index.html") and os.path.exists(_INDEX_HTML):
                    with open(_INDEX_HTML, "rb") as f:
                        return self._send(f.read(), 200, "text/html")
                self._send("not found", 404, "text/plain")

            def do_PUT(self):
                if self.path.rstrip("/") != "/api":
                    return self._send("not found", 404, "text/plain")
                try:
                    n = int(self.headers.get("Content-Length", "0"))
                    req = json.loads(self.rfile.read(n) or b"{}")
                except Exception:
                    return self._send("invalid JSON body", 400, "text/plain")
                body, status = resource.put(req)
                self._send(body, status, "application/json" if isinstance(body, dict) else "text/plain")

            do_POST = do_PUT

            def log_message(self, fmt, *args):
                pass

        self.handler = Handler

    def run(self, url="0.0.0.0", port=5000):
        self.httpd = ThreadingHTTPServer((url, port), self.handler)
        self.httpd.serve_forever()
