"""TransD — host-side descriptor.  TransD.py:18-98: dynamic mapping e + (e . ent_transfer[e]) rel_transfer[r]."""
from .Model import Model


class TransD(Model):
    name = "TransD"
    predict_keepdims = True

    def table_shapes(self):
        c = self.config
        return {"ent_embeddings": (c.entTotal, c.hidden_size), "rel_embeddings": (c.relTotal, c.hidden_size),
                "ent_transfer": (c.entTotal, c.hidden_size), "rel_transfer": (c.relTotal, c.hidden_size)}
