"""TransR — host-side descriptor.  TransR.py:10-87: entities are mapped into relation space by the relation's matrix, e . M_r (M_r = transfer_matrix[r] as [ent_size, rel_size])."""
from .Model import Model


class TransR(Model):
    name = "TransR"
    predict_keepdims = True

    def table_shapes(self):
        c = self.config
        return {"ent_embeddings": (c.entTotal, c.ent_size), "rel_embeddings": (c.relTotal, c.rel_size),
                "transfer_matrix": (c.relTotal, c.ent_size * c.rel_size)}
