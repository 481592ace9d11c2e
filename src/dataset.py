"""Mirror of reference src/dataset.py: `OwlDataset`, `get_dataloaders`, same names, arguments and return values.

With the default `image_processor` the items are what the reference yields (HF processor on the CPU workers).
With `image_processor=None` (`get_dataloaders(device_preprocess=True)`) an item's image is the RAW uint8 RGB array
[H, W, 3]; `OwlViT.forward` recognises uint8 input and runs the resize / rescale / normalise on the device
(`owl_preprocess_image`, Pillow-exact), so the workers only decode.  Like the reference this needs `config.yaml`,
`data/*.json` and the COCO images on disk.
"""
import json
import os
from collections import Counter

import numpy as np
import torch
import yaml
from PIL import Image
from torch.utils.data import DataLoader, Dataset

TRAIN_ANNOTATIONS_FILE = "data/train.json"
TEST_ANNOTATIONS_FILE = "data/test.json"
LABELMAP_FILE = "data/labelmap.json"


def get_images_dir():
    with open("config.yaml", "r") as stream:
        data = yaml.safe_load(stream)["data"]
        return data["images_path"]


class OwlDataset(Dataset):
    def __init__(self, image_processor, annotations_file):
        self.images_dir = get_images_dir()
        self.image_processor = image_processor
        with open(annotations_file) as f:
            data = json.load(f)
            n_total = len(data)
        self.data = [{k: v} for k, v in data.items() if len(v)]
        print(f"Dropping {n_total - len(self.data)} examples due to no annotations")

    def load_image(self, idx: int):
        url = list(self.data[idx].keys()).pop()
        path = os.path.join(self.images_dir, os.path.basename(url))
        return Image.open(path).convert("RGB"), path

    def load_target(self, idx: int):
        annotations = list(self.data[idx].values())
        assert len(annotations) == 1
        annotations = annotations.pop()
        return [a["label"] for a in annotations], [a["bbox"] for a in annotations]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        image, path = self.load_image(idx)
        labels, boxes = self.load_target(idx)
        w, h = image.size
        metadata = {"width": w, "height": h, "impath": path}
        if self.image_processor is None:
            image = torch.from_numpy(np.asarray(image).copy())              # raw uint8 [H, W, 3]
        else:
            image = self.image_processor(images=image, return_tensors="pt")["pixel_values"].squeeze(0)
        return image, torch.tensor(labels), torch.tensor(boxes), metadata


def get_dataloaders(train_annotations_file=TRAIN_ANNOTATIONS_FILE, test_annotations_file=TEST_ANNOTATIONS_FILE,
                    device_preprocess: bool = False):
    if device_preprocess:
        image_processor = None
    else:
        from transformers import OwlViTProcessor
        image_processor = OwlViTProcessor.from_pretrained("google/owlvit-base-patch32")
    train_dataset = OwlDataset(image_processor, train_annotations_file)
    test_dataset = OwlDataset(image_processor, test_annotations_file)
    with open(LABELMAP_FILE) as f:
        labelmap = json.load(f)
    train_labelcounts = Counter()
    for i in range(len(train_dataset)):
        train_labelcounts.update(train_dataset.load_target(i)[0])
    # scales must be in order (reference src/dataset.py:93-98)
    scales = np.array([train_labelcounts[i] for i in sorted(train_labelcounts.keys())])
    scales = (np.round(np.log(scales.max() / scales) + 3, 1)).tolist()
    train_dataloader = DataLoader(train_dataset, batch_size=1, shuffle=True, num_workers=4)
    test_dataloader = DataLoader(test_dataset, batch_size=1, shuffle=False, num_workers=4)
    return train_dataloader, test_dataloader, scales, labelmap
