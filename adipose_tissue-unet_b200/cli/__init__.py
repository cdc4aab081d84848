"""Re-hosted argparse front-ends of the reference scripts (SURVEY.md section 8b "CLI surface"): same flags, same
output files; every array operation goes through libadipose_b200.so.

  python -m adipose_unet_b200.cli.infer     == Segmentation/segmentation_inference.py     (:324-350)
  python -m adipose_unet_b200.cli.recon     == Segmentation/reconstruct_full_images.py    (:882-929)
  python -m adipose_unet_b200.cli.evaluate  == Segmentation/full_evaluation_enhanced.py   (:1989-2036)
  python -m adipose_unet_b200.cli.train     == Segmentation/train_adipose_unet_v3.py      (:1455-1630)
"""
