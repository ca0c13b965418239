"""The reference's hyper-parameters that shape the hot path (config/config.yaml:6,13,93-156), as a plain dict."""
import copy

_DEFAULT = {
    "EXPERIMENT": {"ID": "InsMOS"},
    "DATA": {
        "POINT_CLOUD_RANGE": [-60, -50, -3, 60, 50, 1],
        "CLASE_NAME": ["Car", "Pedestrian", "Cyclist"],
        "TRANSFORM": True,
        "POSES": "poses.txt",
        "SHUFFLE": True,
        "NUM_WORKER": 4,
        "DELTA_T_DATA": 0.1,
        "VOXEL_SIZE": [0.1, 0.1, 0.1],
        "SEMANTIC_CONFIG_FILE": "./config/semantic-kitti-mos.yaml",
        "SPLIT": {"TRAIN": [0, 1, 2, 3, 4, 5, 6, 7, 9, 10], "VAL": [8], "TEST": [8]},
    },
    "TRAIN": {"MAX_EPOCH": 60, "LR": 0.0001, "LR_EPOCH": 1, "LR_DECAY": 0.99, "WEIGHT_DECAY": 0.0001,
              "BATCH_SIZE": 1, "ACC_BATCHES": 1, "AUGMENTATION": True},
    "MODEL": {
        "DELTA_T_PREDICTION": 0.1,
        "N_PAST_STEPS": 10,
        "USE_MOTION_LOSS": True,
        "POINT_FEATURE_ENCODING": {"encoding_type": "absolute_coordinates_encoding",
                                   "used_feature_list": ["x", "y", "z", "intensity"],
                                   "src_feature_list": ["x", "y", "z", "intensity"]},
        "VFE": {"NAME": "MeanVFE"},
        "BACKBONE_3D": {"NAME": "VoxelBackBone8x"},
        "MAP_TO_BEV": {"NAME": "HeightCompression", "NUM_BEV_FEATURES": 256},
        "BACKBONE_2D": {"NAME": "BaseBEVBackbone", "LAYER_NUMS": [5], "LAYER_STRIDES": [1], "NUM_FILTERS": [128],
                        "UPSAMPLE_STRIDES": [2], "NUM_UPSAMPLE_FILTERS": [256]},
        "DENSE_HEAD": {"NAME": "CenterHead", "CLASS_AGNOSTIC": False, "CLASE_NAME": ["Car", "Pedestrian", "Cyclist"],
                       "NUM_CLASS": 3, "USE_DIRECTION_CLASSIFIER": False,
                       "TARGET_ASSIGNER_CONFIG": {"MAX_OBJS": 100, "VOXEL_SIZE": [0.1, 0.1, 0.1], "OUT_SIZE_FACTOR": 4,
                                                  "GAUSSIAN_OVERLAP": 0.1, "MIN_RADIUS": 2, "BOX_CODER": "ResidualCoder"},
                       "LOSS_CONFIG": {"LOSS_WEIGHTS": {"cls_weight": 1.0, "loc_weight": 2.0,
                                                        "code_weights": [1.0] * 8}}},
        "POST_PROCESSING": {"RECALL_THRESH_LIST": [0.3, 0.5, 0.7], "SCORE_THRESH": 0.1, "OUTPUT_RAW_SCORE": False,
                            "EVAL_METRIC": "kitti",
                            "NMS_CONFIG": {"MULTI_CLASSES_NMS": False, "NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.01,
                                           "NMS_PRE_MAXSIZE": 4096, "NMS_POST_MAXSIZE": 500}},
    },
}

# config/semantic-kitti-mos.yaml: learning_map_inv has 3 entries (0 unlabeled, 1 static, 2 moving), class 0 ignored
SEMANTIC_DEFAULT = {"learning_map_inv": {0: 0, 1: 9, 2: 251}, "learning_ignore": {0: True, 1: False, 2: False}}


def default_config():
    return copy.deepcopy(_DEFAULT)
