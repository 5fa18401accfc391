"""abstracts-search_b200 — B200-native (sm_100a) implementation of the abstracts-search hot path:
stella_en_1.5B_v5 bulk embedding (SentenceTransformer.encode surface) and faiss-style IVF/Flat
inner-product search (faiss.Index train/add/search surface), over libabsb200.so.

The directory name carries a hyphen (it is the repo's required layout); import it with
`importlib.import_module("abstracts-search_b200")` or through the `abstracts_search_b200` alias
module at the repo root.
"""
from ._lib import AbsbError, LIB_PATH, build, header_functions, lib  # noqa: F401
from .index import (METRIC_INNER_PRODUCT, METRIC_L2, ClusteringParameters, IndexFlatIP,  # noqa: F401
                    IndexIVFFlat, ParameterSpace, SearchParametersIVF, extract_index_ivf,
                    index_factory, read_index, write_index)
from .sharded import DeviceOps, ShardedIndexIVFFlat, merge_partials_host, owner_of_list  # noqa: F401
from .encoder import STELLA_1_5B, Encoder, EncoderConfig, SentenceTransformer  # noqa: F401
from .peer import PeerExchange  # noqa: F401
from .pipeline import QueryPipeline  # noqa: F401
from . import faiss_io, oa_jsonl, peer, pipeline, store, synth, tune  # noqa: F401
