from .tokenizer import build_tokenizer
