"""Host-side mirror of the reference's ROI-head interfaces for the region-scoring path."""
from .poolers import ROIPooler, convert_boxes_to_pooler_format
from .class_heads import OpenVocabularyClassifier
from .roi_heads import (InstanceRefinementOutputLayers, ObjectMiningOutputLayers, fast_rcnn_inference,
                        fast_rcnn_inference_single_image, get_image_level_gt, get_pgt_top_k,
                        label_proposals_wsl)
from .proposal_utils import find_top_rpn_proposals, find_top_rpn_proposals_group
from .wsovod_heads import WSOVODMixedDatasetsROIHeads, WSOVODROIHeads, get_pgt_mist

__all__ = ["ROIPooler", "convert_boxes_to_pooler_format", "OpenVocabularyClassifier", "ObjectMiningOutputLayers",
           "InstanceRefinementOutputLayers", "fast_rcnn_inference", "fast_rcnn_inference_single_image",
           "get_image_level_gt", "get_pgt_top_k", "label_proposals_wsl", "find_top_rpn_proposals", "find_top_rpn_proposals_group", "WSOVODROIHeads",
           "WSOVODMixedDatasetsROIHeads", "get_pgt_mist"]
