"""TransH — host-side descriptor.  TransH.py:7-82: entities are projected on the relation's hyperplane, e - (e.n)n with n = l2n(normal_vectors[r])."""
from .Model import Model


class TransH(Model):
    name = "TransH"
    predict_keepdims = True

    def table_shapes(self):
        c = self.config
        return {"ent_embeddings": (c.entTotal, c.hidden_size), "rel_embeddings": (c.relTotal, c.hidden_size),
                "normal_vectors": (c.relTotal, c.hidden_size)}
