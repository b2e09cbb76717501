"""Golden vectors for the dataset-preprocessing mirror: the UNMODIFIED reference GeneralDataset
(gelslim_depth/datasets/general_dataset.py) is instantiated on a temporary directory holding one synthetic object
file; every normalised sample it serves is recorded next to the raw tensors.  Run in the build container:
    python tests/golden/make_dataset_golden.py        (reads /root/reference, writes tests/golden/dataset_preprocess.pt)"""
import os, sys, tempfile
import torch

sys.path.insert(0, "/root/reference")
from gelslim_depth.datasets.general_dataset import GeneralDataset  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(31)
    n, h, w = 3, 40, 54
    data = {"tactile_image": torch.randint(0, 256, (n, 6, h, w), generator=g).float(),
            "base_tactile_image": torch.randint(0, 256, (1, 6, h, w), generator=g).float(),
            "depth_image": -2.0 * torch.rand(n, 2, h, w, generator=g)}
    cases = {}
    with tempfile.TemporaryDirectory() as d:
        torch.save(data, os.path.join(d, "obj0.pt"))
        cfgs = {
            "shipped": dict(use_difference_image=True, depth_normalization_method="min_max_to_0_-1", image_normalization_method="0_255_to_0_1",
                            separate_fingers=True, downsample_factor=0.5, depth_image_blur_kernel=1,
                            depth_normalization_parameters=(-1.9180814027786255, 0.0), norm_scale=0.9, interp_method="area"),
            "blur_meanstd": dict(use_difference_image=True, depth_normalization_method="min_max_to_-1_1", image_normalization_method="mean_std",
                                 separate_fingers=True, downsample_factor=0.5, depth_image_blur_kernel=5, norm_scale=1.0, interp_method="area"),
            "joint_nodiff": dict(use_difference_image=False, depth_normalization_method="mean_std", image_normalization_method="0_255_to_-1_1",
                                 separate_fingers=False, downsample_factor=0.5, depth_image_blur_kernel=3, norm_scale=1.0, interp_method="area"),
        }
        for name, kw in cfgs.items():
            ds = GeneralDataset(directory=d, pt_file_list=["obj0.pt"], **kw)
            xs = torch.stack([ds[i]["tactile_image"] for i in range(len(ds))])
            ys = torch.stack([ds[i]["depth_image"] for i in range(len(ds))])
            cases[name] = {"kwargs": kw, "input_tactile_image_size": tuple(ds.input_tactile_image_size),
                           "image_normalization_parameters": ds.image_normalization_parameters,
                           "depth_normalization_parameters": ds.depth_normalization_parameters,
                           "tactile_image": xs, "depth_image": ys}
            print(name, xs.shape, ys.shape, float(xs.mean()), float(ys.mean()))
    torch.save({"data": data, "cases": cases}, os.path.join(HERE, "dataset_preprocess.pt"))


if __name__ == "__main__":
    main()
