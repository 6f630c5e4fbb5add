"""Data pipeline: mmap indexed datasets, GPT / instruction / BERT / T5 / ICT datasets, samplers, C++ index builders."""
