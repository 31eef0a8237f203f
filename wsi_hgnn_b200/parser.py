"""`parse_gnn_model` for the three model families on the hot path - the mirror of reference parser.py:48-174 for
`GNN.name` in {"HGT", "HEAT2", "HEAT4"}: same config keys, same node_dict / edge_dict construction (edge_dict order:
for r in edge_types, for s, for t - parser.py:125-134), same constructor arguments (note `dropuout`, parser.py:155,169).
Any other name raises NotImplementedError like the reference's final branch (the other models are outside this path)."""
from typing import Dict

from .models import HEATNet2, HEATNet4, HGT


def node_and_edge_dicts(n_node_types: int, edge_types):
    """(node_dict, edge_dict) exactly as parser.py:125-134 builds them."""
    canonical_etypes = [(str(s), r, str(t)) for r in edge_types for s in range(n_node_types) for t in range(n_node_types)]
    node_dict = {str(i): i for i in range(n_node_types)}
    edge_dict = {et: i for i, et in enumerate(canonical_etypes)}
    return node_dict, edge_dict


def parse_gnn_model(config_gnn: Dict):
    gnn_name = config_gnn["name"]
    if gnn_name == "HGT":                                                      # parser.py:124-143
        node_dict, edge_dict = node_and_edge_dicts(config_gnn["n_node_types"], config_gnn["edge_types"])
        return HGT(node_dict, edge_dict, in_dim=config_gnn["in_dim"], hidden_dim=config_gnn["hidden_dim"],
                   out_dim=config_gnn["out_dim"], n_layers=config_gnn["num_layers"], n_heads=config_gnn["num_heads"])
    if gnn_name in ("HEAT2", "HEAT4"):                                         # parser.py:145-171
        node_dict = {str(i): i for i in range(config_gnn["n_node_types"])}
        cls = HEATNet2 if gnn_name == "HEAT2" else HEATNet4
        return cls(in_dim=config_gnn["in_dim"], hidden_dim=config_gnn["hidden_dim"], out_dim=config_gnn["out_dim"],
                   n_layers=config_gnn["num_layers"], n_heads=config_gnn["n_heads"], node_dict=node_dict,
                   dropuout=config_gnn["feat_drop"], graph_pooling_type=config_gnn["graph_pooling_type"])
    raise NotImplementedError("This GNN model is not implemented")
