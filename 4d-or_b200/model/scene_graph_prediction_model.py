"""``SGPNModelWrapper`` -- the model API the reference's ``main.py`` drives
(SGH/model/scene_graph_prediction_model.py:30-242; constructed at SGP/main.py:58-59).

Same constructor arguments, sub-module attribute names (``obj_encoder``, ``rel_encoder``, ``gcn``,
``obj_predictor``, ``rel_predictor``, ``full_image_feature_reduction``) and therefore the same 188
``state_dict`` keys; same ``forward(batch, return_meta_data)`` results; ``training_step`` /
``validation_step`` return the reference loss ``lambda_o * nll(obj) + nll(rel)`` (:139-141).

It is a plain ``nn.Module`` (pytorch-lightning is a trainer dependency of the reference, not part of
the hot path); a Lightning trainer can still drive it because the step / optimizer hooks keep their
names.  Host-side metric bookkeeping (sklearn reports, :182-238) is outside the hot path.

Batches may hold ONE scene (the reference's ``batch_size=1``) or several concatenated scenes
(``edge_indices`` already offset per scene); BatchNorm statistics are then pooled over the batch,
which is exactly what the reference modules compute when handed the same concatenated tensors.

Image branch (``IMAGE_INPUT == 'full'``): the timm EfficientNet feature extractor is third-party and
outside the hot path; the model consumes its OUTPUT features ``batch['full_image_features']``
((6, 2048) per scene, or (S, 6, 2048) with ``batch['edge_scene']`` (E,) scene ids) and applies the
reference's ``full_image_feature_reduction`` + flatten + late fusion (:98-102).
"""
import torch
import torch.nn.functional as F
import torch.optim as optim
from torch import nn

from .. import dense
from .network_PointNet import PointNetCls, PointNetRelCls
from .network_PointNet2 import PointNetfeat
from .network_TripletGCN import TripletGCNModel

IMAGE_NUM_FEATURES = 2048  # tf_efficientnet_b5_ns.num_features (model_utils.py:10-22)


class SGPNModelWrapper(nn.Module):
    def __init__(self, config, num_class, num_rel, weights_obj, weights_rel, relationNames):
        super().__init__()
        self.config = config
        self.mconfig = config['MODEL']
        self.n_object_types = 6
        self.weights_obj = weights_obj
        self.weights_rel = weights_rel
        self.relationNames = relationNames
        self.lr = float(self.config['LR'])

        self.obj_encoder = PointNetfeat(input_dim=6, out_size=self.mconfig['point_feature_size'],
                                        input_dropout=self.mconfig['INPUT_DROPOUT'])
        self.rel_encoder = PointNetfeat(input_dim=7, out_size=self.mconfig['edge_feature_size'],
                                        input_dropout=self.mconfig['INPUT_DROPOUT'])
        self.use_image = self.config['IMAGE_INPUT'] == 'full'
        if self.use_image:
            self.full_image_feature_reduction = nn.Linear(
                IMAGE_NUM_FEATURES, self.mconfig['FULL_IMAGE_EMBEDDING_SIZE'] // 6)

        self.gcn = TripletGCNModel(num_layers=self.mconfig['N_LAYERS'],
                                   dim_node=self.mconfig['point_feature_size'],
                                   dim_edge=self.mconfig['edge_feature_size'],
                                   dim_hidden=self.mconfig['gcn_hidden_feature_size'])
        self.obj_predictor = PointNetCls(num_class, in_size=self.mconfig['point_feature_size'],
                                         batch_norm=False, drop_out=True)
        self.rel_predictor = PointNetRelCls(
            num_rel, in_size=self.mconfig['edge_feature_size'], batch_norm=False, drop_out=True,
            image_embedding_size=self.mconfig['FULL_IMAGE_EMBEDDING_SIZE'] if self.use_image else None,
            n_object_types=self.n_object_types)

    # ------------------------------------------------------------------ forward (reference :87-109)
    # the two encoders are independent until the GCN: the (small) object encoder runs on a side stream so that
    # its latency-bound kernels (12 clouds per scene) share the GPU with the edge encoder's (66 clouds per scene)
    overlap_encoders = True

    def _encode(self, batch):
        obj_points, rel_points = batch['obj_points'], batch['rel_points']
        if not (self.overlap_encoders and obj_points.is_cuda):
            return self.obj_encoder(obj_points), self.rel_encoder(rel_points)
        main = torch.cuda.current_stream(obj_points.device)
        side = self.__dict__.get('_side_stream')
        if side is None or side.device != obj_points.device:
            side = torch.cuda.Stream(device=obj_points.device, priority=-1)     # same priority as a graphed step's capture stream
            self.__dict__['_side_stream'] = side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            obj_feature = self.obj_encoder(obj_points)
        rel_feature = self.rel_encoder(rel_points)
        main.wait_stream(side)
        obj_feature.record_stream(main)
        return obj_feature, rel_feature

    def forward(self, batch, return_meta_data=False):
        obj_feature, rel_feature = self._encode(batch)
        gcn_obj_feature, gcn_rel_feature = self.gcn(obj_feature, rel_feature, batch['edge_indices'])

        obj_cls = self.obj_predictor(gcn_obj_feature if self.mconfig['OBJ_PRED_FROM_GCN'] else obj_feature)
        image_embeddings = None
        if self.use_image:
            fi = batch['full_image_features']
            feats = dense.linear(fi.reshape(-1, fi.shape[-1]), self.full_image_feature_reduction).reshape(*fi.shape[:-1], -1)
            if feats.dim() == 2:            # one scene: (6, 128) -> (768,)
                image_embeddings = feats.flatten()
            else:                           # S scenes: (S, 6, 128) -> per-edge (E, 768)
                image_embeddings = feats.flatten(1).index_select(0, batch['edge_scene'])
        rel_cls = self.rel_predictor(gcn_rel_feature,
                                     relation_objects_one_hot=batch['relation_objects_one_hot'],
                                     image_embeddings=image_embeddings)
        if return_meta_data:
            return obj_cls, rel_cls, obj_feature, rel_feature, gcn_obj_feature, gcn_rel_feature, None
        return obj_cls, rel_cls

    # ------------------------------------------------------------------ steps (reference :134-177)
    def _class_weights(self, device):
        """the class-weight vectors on `device`, uploaded once (the reference re-uploads them every step, :139-140; a copy
        from pageable host memory can also not be captured into a CUDA graph)"""
        cache = self.__dict__.setdefault('_weights_on', {})
        key = (str(device), id(self.weights_obj), id(self.weights_rel))
        if key not in cache:
            cache.clear()
            cache[key] = (self.weights_obj.to(device), self.weights_rel.to(device))
        return cache[key]

    def loss(self, obj_pred, rel_pred, batch):
        w_obj, w_rel = self._class_weights(obj_pred.device)
        loss_obj = F.nll_loss(obj_pred, batch['gt_class'], weight=w_obj)
        loss_rel = F.nll_loss(rel_pred, batch['gt_rels'], weight=w_rel)
        return self.mconfig['lambda_o'] * loss_obj + loss_rel

    # per-take relation predictions for the epoch metrics (reference :113-132); set by the trainer (sg4d.metrics.RelationMetrics)
    metrics = None

    def update_metrics(self, batch, rel_pred, split='train'):
        if self.metrics is not None:
            self.metrics.update(batch, rel_pred, split)

    def reset_metrics(self, split=None):
        if self.metrics is not None:
            self.metrics.reset(split)

    def training_step(self, batch, batch_idx=0):
        obj_pred, rel_pred = self(batch)
        self.update_metrics(batch, rel_pred, 'train')
        return self.loss(obj_pred, rel_pred, batch)

    def validation_step(self, batch, batch_idx=0):
        obj_pred, rel_pred = self(batch)
        self.update_metrics(batch, rel_pred, 'val')
        return self.loss(obj_pred, rel_pred, batch)

    def predict_step(self, batch, batch_idx=0, dataloader_idx=0):
        _, rel_pred = self(batch)
        predicted = rel_pred.detach().argmax(1).tolist()
        none_id = self.relationNames.index('none')
        edges = batch['edge_indices'].t().tolist()
        relations = []
        for (start, end), rel in zip(edges, predicted):
            if rel == none_id:
                continue
            relations.append((batch['objs_json'][start + 1], self.relationNames[rel],
                              batch['objs_json'][end + 1]))
        return batch['scan_id'], relations

    def configure_optimizers(self):
        return optim.AdamW(params=self.parameters(), lr=self.lr, weight_decay=float(self.config['W_DECAY']))
