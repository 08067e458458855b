"""TransE — host-side descriptor.  TransE.py:5-57: score = |l2n(h) + l2n(r) - l2n(t)|_1; predict = mean over d -> [N]."""
from .Model import Model


class TransE(Model):
    name = "TransE"
    predict_keepdims = False

    def table_shapes(self):
        c = self.config
        return {"ent_embeddings": (c.entTotal, c.hidden_size), "rel_embeddings": (c.relTotal, c.hidden_size)}
