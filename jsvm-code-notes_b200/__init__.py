"""B200-native JSVM motion-estimation hot path (full search + quarter-pel refinement + half-pel planes).

Importing the package binds the in-tree CUDA library (libjsvm_me_b200.so); a missing library is a
hard error -- there is no CPU implementation of any compute entry point in this package.
"""
from . import abi
from .abi import (PlanInfo, MC_JOB_DTYPE, PIC_TASK_DTYPE, PIC_MB_DTYPE, MB_MOTION_DTYPE, ResizeParams, MotionUpsampleParams,
                  JOB_DTYPE, RESULT_DTYPE, MODE_16x16, MODE_16x8, MODE_8x16, BLK_8x8, BLK_8x4, BLK_4x8, BLK_4x4,
                  DF_SAD, DF_HADAMARD, load_library)
from .me import (MotionEstimator, MotionEstimationError, MotionEstimationQuarterPel, QuarterPelFilter, Frame,
                 cost_factor_sad, BLOCK_SIZE, plan_host, GROUP_DTYPE)

load_library()
