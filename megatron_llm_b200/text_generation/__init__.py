"""Inference: incremental decoding with a (native-GQA) KV cache, sampling, beam search, scoring, REST server."""
from .api import beam_search_and_post_process, generate, generate_and_post_process  # noqa: F401
