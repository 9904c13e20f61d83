"""torch_geometric.data.Data stand-in (itexperiments.py:23,188-189,199,258,315; rd2pd.py:9,127)."""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getattr__(self, name):                      # only reached for attributes never set
        raise AttributeError(name)

    @property
    def num_nodes(self):
        x = self.__dict__.get("x")
        if x is not None:
            return x.size(0)
        ei = self.__dict__.get("edge_index")
        return int(ei.max()) + 1 if ei is not None and ei.numel() else 0

    @property
    def num_node_features(self):
        x = self.__dict__.get("x")
        if x is None:
            return 0
        return 1 if x.dim() == 1 else x.size(1)

    num_features = num_node_features

    @property
    def num_edges(self):
        ei = self.__dict__.get("edge_index")
        return 0 if ei is None else ei.size(1)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else v
        return out

    def to(self, device, *args, **kwargs):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device, *args, **kwargs)
        return self

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def __repr__(self):
        parts = [f"{k}={list(v.shape) if torch.is_tensor(v) else v}" for k, v in self.__dict__.items()
                 if v is not None]
        return f"Data({', '.join(parts)})"
