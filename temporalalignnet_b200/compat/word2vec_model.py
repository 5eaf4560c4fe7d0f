"""`from word2vec_model import Word2VecModel` (model/tan_model.py:8) resolves here when this directory is on sys.path."""
from temporalalignnet_b200.word2vec_model import Word2VecModel  # noqa: F401
