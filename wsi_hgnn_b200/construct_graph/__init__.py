from .graph_constructor import GraphConstructor, Hnsw, construct_graph_arrays

__all__ = ["GraphConstructor", "Hnsw", "construct_graph_arrays"]
