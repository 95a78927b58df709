def is_xformers_available():
    """xformers is not installed on the build or GPU boxes -> the reference takes its
    naive ``get_attention_scores`` + ``bmm`` branch (attention.py:156-158)."""
    return False
