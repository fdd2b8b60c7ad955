"""``faiss`` alias: lets the reference's scripts (`import faiss`, cath/search.py:5,
pfam/proteins_search.py:4, seqvec_search/main.py:9) bind to knn_b200 unchanged when
``knn-for-homology_b200/`` is on sys.path.  Only the flat-search surface exists."""
from knn_b200 import *  # noqa: F401,F403
from knn_b200 import __all__  # noqa: F401
