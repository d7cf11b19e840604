#!/usr/bin/env python3
"""Kernel-path sweep on one B200: forward and backward of one MMTM block timed separately
through the C ABI (CUDA events around each call, L2 flushed between iterations), for every
shape x batch x kernel-path variant.  Writes gpurun_out/sweep.json and prints a table.

    python scripts/sweep.py [--batches 32,256,1024] [--iters 12] [--variants ...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import SHAPES, BlockBuffers, measured_peak  # noqa: E402
from greedy_multimodal_learning_b200 import _lib as L  # noqa: E402

VARIANTS = {
    # name: (flags, tunables)
    "auto": (0, {}),
    "old": (0, {"tile_kind": 2}),
    "noswitch": (0, {"tile_switch": 0}),
    "x_nodeps1": (L.F_FORCE_TILE, {"tile_nodeps": 1}),
    "x_nodeps2": (L.F_FORCE_TILE, {"tile_nodeps": 2}),
    "x_nodeps3": (L.F_FORCE_TILE, {"tile_nodeps": 3}),
    "x_nodeps4": (L.F_FORCE_TILE, {"tile_nodeps": 4}),
    "x_nodeps1_g3": (L.F_FORCE_TILE, {"tile_nodeps": 1, "tile_gemm_ctas": 3}),
    "x_nodeps2_g3": (L.F_FORCE_TILE, {"tile_nodeps": 2, "tile_gemm_ctas": 3}),
    "x_nodeps4_g3": (L.F_FORCE_TILE, {"tile_nodeps": 4, "tile_gemm_ctas": 3}),
    "k_ks8_g72": (L.F_FORCE_TILE, {"tile_ksplit_tiles": 8, "tile_gemm_ctas": 72}),
    "k_ks8_g48": (L.F_FORCE_TILE, {"tile_ksplit_tiles": 8, "tile_gemm_ctas": 48}),
    "k_ks16_g72": (L.F_FORCE_TILE, {"tile_ksplit_tiles": 16, "tile_gemm_ctas": 72}),
    "k_ks4_g72": (L.F_FORCE_TILE, {"tile_ksplit_tiles": 4, "tile_gemm_ctas": 72}),
    "k_ks8_g72_l3": (L.F_FORCE_TILE, {"tile_ksplit_tiles": 8, "tile_gemm_ctas": 72, "tile_lag": 3}),
    "k_ks8_g72_l4": (L.F_FORCE_TILE, {"tile_ksplit_tiles": 8, "tile_gemm_ctas": 72, "tile_lag": 4}),
    "k_g72_l4": (L.F_FORCE_TILE, {"tile_gemm_ctas": 72, "tile_lag": 4}),
    "k_g72_l3": (L.F_FORCE_TILE, {"tile_gemm_ctas": 72, "tile_lag": 3}),
    "k_g72_m64": (L.F_FORCE_TILE, {"tile_gemm_ctas": 72, "tile_m": 64}),
    "k_ks8_g72_m64": (L.F_FORCE_TILE, {"tile_ksplit_tiles": 8, "tile_gemm_ctas": 72, "tile_m": 64}),
    "x_ronly": (L.F_FORCE_TILE, {"tile_nodeps": 5}),
    "x_sonly": (L.F_FORCE_TILE, {"tile_nodeps": 6}),
    "x_ronly_c14": (L.F_FORCE_TILE, {"tile_nodeps": 5, "tile_chunk_kb": 14}),
    "x_sonly_c14": (L.F_FORCE_TILE, {"tile_nodeps": 6, "tile_chunk_kb": 14}),
    "x_ronly_c56": (L.F_FORCE_TILE, {"tile_nodeps": 5, "tile_chunk_kb": 56}),
    "p_ronly_p1": (L.F_FORCE_TILE, {"tile_nodeps": 5, "tile_rpol": 1}),
    "p_ronly_p2": (L.F_FORCE_TILE, {"tile_nodeps": 5, "tile_rpol": 2}),
    "p_tile_p1": (L.F_FORCE_TILE, {"tile_rpol": 1}),
    "p_tile_p2": (L.F_FORCE_TILE, {"tile_rpol": 2}),
    "q_rload": (L.F_FORCE_TILE, {"tile_nodeps": 7}),
    "q_sload": (L.F_FORCE_TILE, {"tile_nodeps": 8}),
    "q_rload_c56": (L.F_FORCE_TILE, {"tile_nodeps": 7, "tile_chunk_kb": 56}),
    "q_rload_c14": (L.F_FORCE_TILE, {"tile_nodeps": 7, "tile_chunk_kb": 14}),
    "s_slots4": (L.F_FORCE_TILE, {"tile_max_slots": 4}),
    "d_draw2": (L.F_FORCE_TILE, {"tile_draw": 2}),
    "d_draw8": (L.F_FORCE_TILE, {"tile_draw": 8}),
    "d_draw16": (L.F_FORCE_TILE, {"tile_draw": 16}),
    "s_slots2": (L.F_FORCE_TILE, {"tile_max_slots": 2}),
    "s_slots4_c14": (L.F_FORCE_TILE, {"tile_max_slots": 4, "tile_chunk_kb": 14}),
    "s_slots8_c14": (L.F_FORCE_TILE, {"tile_max_slots": 8, "tile_chunk_kb": 14}),
    "s_rload_slots4": (L.F_FORCE_TILE, {"tile_max_slots": 4, "tile_nodeps": 7}),
    "s_rload_slots2": (L.F_FORCE_TILE, {"tile_max_slots": 2, "tile_nodeps": 7}),
    "c_rload_x2": (L.F_FORCE_TILE, {"tile_nodeps": 7, "tile_split_copies": 2}),
    "c_rload_x4": (L.F_FORCE_TILE, {"tile_nodeps": 7, "tile_split_copies": 4}),
    "c_rload_x7": (L.F_FORCE_TILE, {"tile_nodeps": 7, "tile_split_copies": 7}),
    "c_rload_c56_x2": (L.F_FORCE_TILE, {"tile_nodeps": 7, "tile_split_copies": 2, "tile_chunk_kb": 56}),
    "c_rload_c112": (L.F_FORCE_TILE, {"tile_nodeps": 7, "tile_chunk_kb": 100}),
    "d_draw1": (L.F_FORCE_TILE, {"tile_draw": 1}),
    "d_draw2": (L.F_FORCE_TILE, {"tile_draw": 2}),
    "d_draw4": (L.F_FORCE_TILE, {"tile_draw": 4}),
    "d_draw8": (L.F_FORCE_TILE, {"tile_draw": 8}),
    "d_draw16": (L.F_FORCE_TILE, {"tile_draw": 16}),
    "d_c56": (L.F_FORCE_TILE, {"tile_chunk_kb": 56}),
    "d_c40": (L.F_FORCE_TILE, {"tile_chunk_kb": 40}),
    "f_c28": (L.F_FORCE_TILE, {"tile_chunk_kb_fwd": 28}),
    "f_c56": (L.F_FORCE_TILE, {"tile_chunk_kb_fwd": 56}),
    "f_c100": (L.F_FORCE_TILE, {"tile_chunk_kb_fwd": 100}),
    "l_cs8": (0, {"tile_kind": 2, "fused_cluster": 8}),
    "l_cs4": (0, {"tile_kind": 2, "fused_cluster": 4}),
    "l_occ5": (0, {"tile_kind": 2, "fused_occ": 5}),
    "l_stash0": (0, {"tile_kind": 2, "fused_stash_kb": 0}),
    "l_stash12": (0, {"tile_kind": 2, "fused_stash_kb": 12}),
    "l_stash40": (0, {"tile_kind": 2, "fused_stash_kb": 40}),
    "l_cs8_stash40": (0, {"tile_kind": 2, "fused_cluster": 8, "fused_stash_kb": 40}),
    "l_cs8_occ5": (0, {"tile_kind": 2, "fused_cluster": 8, "fused_occ": 5}),
    "g_lag5": (L.F_FORCE_TILE, {"tile_lag": 5}),
    "g_lag6": (L.F_FORCE_TILE, {"tile_lag": 6}),
    "g_lag7": (L.F_FORCE_TILE, {"tile_lag": 7}),
    "g_lag8": (L.F_FORCE_TILE, {"tile_lag": 8}),
    "g_lag8_m64": (L.F_FORCE_TILE, {"tile_lag": 8, "tile_m": 64}),
    "g_lag16_m64": (L.F_FORCE_TILE, {"tile_lag": 16, "tile_m": 64}),
    "g_lag12_m64": (L.F_FORCE_TILE, {"tile_lag": 12, "tile_m": 64}),
    "old_ws0": (0, {"tile_kind": 2, "fused_wsmem": 0}),
    "old_ws1": (0, {"tile_kind": 2, "fused_wsmem": 1}),
    "old_ws2": (0, {"tile_kind": 2, "fused_wsmem": 2}),
    "old_ws3": (0, {"tile_kind": 2, "fused_wsmem": 3}),
    "old_ws3_st0": (0, {"tile_kind": 2, "fused_wsmem": 3, "fused_stash_kb": 0}),
    "old_ws3_st12": (0, {"tile_kind": 2, "fused_wsmem": 3, "fused_stash_kb": 12}),
    "old_ws3_st8": (0, {"tile_kind": 2, "fused_wsmem": 3, "fused_stash_kb": 8}),
    "old_ws3_st16": (0, {"tile_kind": 2, "fused_wsmem": 3, "fused_stash_kb": 16}),
    "old_ws1_st12": (0, {"tile_kind": 2, "fused_wsmem": 1, "fused_stash_kb": 12}),
    "old_ws2_st12": (0, {"tile_kind": 2, "fused_wsmem": 2, "fused_stash_kb": 12}),
    "old_ws1_st0": (0, {"tile_kind": 2, "fused_wsmem": 1, "fused_stash_kb": 0}),
    "old_ws2_st0": (0, {"tile_kind": 2, "fused_wsmem": 2, "fused_stash_kb": 0}),
    "old_ws0_st12": (0, {"tile_kind": 2, "fused_wsmem": 0, "fused_stash_kb": 12}),
    "old_cs8_ws0": (0, {"tile_kind": 2, "fused_cluster": 8, "fused_wsmem": 0}),
    "old_cs8_ws3_st12": (0, {"tile_kind": 2, "fused_cluster": 8, "fused_wsmem": 3, "fused_stash_kb": 12}),
    "old_cs8_ws3_st8": (0, {"tile_kind": 2, "fused_cluster": 8, "fused_wsmem": 3, "fused_stash_kb": 8}),
    "old_cs8_ws3_st0": (0, {"tile_kind": 2, "fused_cluster": 8, "fused_wsmem": 3, "fused_stash_kb": 0}),
    "light_tile": (0, {"tile_light_fwd": 1}),
    "hw_generic": (0, {"fused_hw_special": 0}),
    "c16_ws3": (0, {"tile_kind": 2, "fused_cluster": 16, "fused_wsmem": 3, "fused_stash_kb": 0}),
    "c16_ws3_st12": (0, {"tile_kind": 2, "fused_cluster": 16, "fused_wsmem": 3, "fused_stash_kb": 12}),
    "c16_ws0": (0, {"tile_kind": 2, "fused_cluster": 16, "fused_wsmem": 0}),
    "c16_ws3_g1": (0, {"tile_kind": 2, "fused_cluster": 16, "fused_wsmem": 3, "fused_stash_kb": 0, "fused_group_kb": 1}),
    "p_st0": (0, {"fused_stash_kb": 0}),
    "p_st16": (0, {"fused_stash_kb": 16}),
    "p_occ5": (0, {"fused_occ": 5}),
    "w_nofold": (0, {"tile_wgrad": 0}),
    "w_g16": (0, {"tile_gemm_ctas": 16}),
    "w_g20": (0, {"tile_gemm_ctas": 20}),
    "w_g24": (0, {"tile_gemm_ctas": 24}),
    "w_g32": (0, {"tile_gemm_ctas": 32}),
    "w_g40": (0, {"tile_gemm_ctas": 40}),
    "w_g56": (0, {"tile_gemm_ctas": 56}),
    "w_g64": (0, {"tile_gemm_ctas": 64}),
    "w_g72": (0, {"tile_gemm_ctas": 72}),
    "x_c14": (L.F_FORCE_TILE, {"tile_chunk_kb": 14}),
    "x_c56": (L.F_FORCE_TILE, {"tile_chunk_kb": 56}),
    "sw_g24": (L.F_FORCE_TILE, {"tile_gemm_ctas": 24}),
    "sw_g56": (L.F_FORCE_TILE, {"tile_gemm_ctas": 56}),
    "sw_g72": (L.F_FORCE_TILE, {"tile_gemm_ctas": 72}),
    "sw_l3_g72": (L.F_FORCE_TILE, {"tile_gemm_ctas": 72, "tile_lag": 3}),
    "sw_l4_g56": (L.F_FORCE_TILE, {"tile_gemm_ctas": 56, "tile_lag": 4}),
    "sw_l4_g24": (L.F_FORCE_TILE, {"tile_gemm_ctas": 24, "tile_lag": 4}),
    "sw_l3_g40": (L.F_FORCE_TILE, {"tile_gemm_ctas": 40, "tile_lag": 3}),
    "sw_tile": (L.F_FORCE_TILE, {}),
    "tile": (L.F_FORCE_TILE, {}),
    "tile_lag1": (L.F_FORCE_TILE, {"tile_lag": 1}),
    "tile_lag3": (L.F_FORCE_TILE, {"tile_lag": 3}),
    "tile_lag4": (L.F_FORCE_TILE, {"tile_lag": 4}),
    "tile_m16": (L.F_FORCE_TILE, {"tile_m": 16}),
    "tile_m32": (L.F_FORCE_TILE, {"tile_m": 32}),
    "tile_m64": (L.F_FORCE_TILE, {"tile_m": 64}),
    "tile_m128": (L.F_FORCE_TILE, {"tile_m": 128}),
    "tile_g4": (L.F_FORCE_TILE, {"tile_gemm_ctas": 4}),
    "tile_g8": (L.F_FORCE_TILE, {"tile_gemm_ctas": 8}),
    "tile_g12": (L.F_FORCE_TILE, {"tile_gemm_ctas": 12}),
    "tile_g16": (L.F_FORCE_TILE, {"tile_gemm_ctas": 16}),
    "tile_g24": (L.F_FORCE_TILE, {"tile_gemm_ctas": 24}),
    "tile_g32": (L.F_FORCE_TILE, {"tile_gemm_ctas": 32}),
    "tile_g40": (L.F_FORCE_TILE, {"tile_gemm_ctas": 40}),
    "tile_c14": (L.F_FORCE_TILE, {"tile_chunk_kb": 14}),
    "tile_nodeps": (L.F_FORCE_TILE, {"tile_nodeps": 1, "tile_gemm_ctas": 8}),
    "g_m4_l1": (L.F_FORCE_TILE, {"tile_m": 4, "tile_lag": 1, "tile_gemm_ctas": 8}),
    "g_m4_l2": (L.F_FORCE_TILE, {"tile_m": 4, "tile_lag": 2, "tile_gemm_ctas": 8}),
    "g_m4_l3": (L.F_FORCE_TILE, {"tile_m": 4, "tile_lag": 3, "tile_gemm_ctas": 8}),
    "g_m4_l4": (L.F_FORCE_TILE, {"tile_m": 4, "tile_lag": 4, "tile_gemm_ctas": 8}),
    "g_m4_l6": (L.F_FORCE_TILE, {"tile_m": 4, "tile_lag": 6, "tile_gemm_ctas": 8}),
    "g_m8_l1": (L.F_FORCE_TILE, {"tile_m": 8, "tile_lag": 1, "tile_gemm_ctas": 8}),
    "g_m8_l2": (L.F_FORCE_TILE, {"tile_m": 8, "tile_lag": 2, "tile_gemm_ctas": 8}),
    "g_m8_l3": (L.F_FORCE_TILE, {"tile_m": 8, "tile_lag": 3, "tile_gemm_ctas": 8}),
    "g_m8_l4": (L.F_FORCE_TILE, {"tile_m": 8, "tile_lag": 4, "tile_gemm_ctas": 8}),
    "g_m8_l6": (L.F_FORCE_TILE, {"tile_m": 8, "tile_lag": 6, "tile_gemm_ctas": 8}),
    "g_m16_l1": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 1, "tile_gemm_ctas": 8}),
    "g_m16_l2": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 2, "tile_gemm_ctas": 8}),
    "g_m16_l3": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 3, "tile_gemm_ctas": 8}),
    "g_m16_l4": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 4, "tile_gemm_ctas": 8}),
    "g_m16_l6": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 6, "tile_gemm_ctas": 8}),
    "g_m32_l1": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 1, "tile_gemm_ctas": 8}),
    "g_m32_l2": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 2, "tile_gemm_ctas": 8}),
    "g_m32_l3": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 3, "tile_gemm_ctas": 8}),
    "g_m32_l4": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 4, "tile_gemm_ctas": 8}),
    "g_m32_l6": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 6, "tile_gemm_ctas": 8}),
    "s_m64_l2_g12": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 2, "tile_gemm_ctas": 12}),
    "s_m64_l2_g24": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 2, "tile_gemm_ctas": 24}),
    "s_m64_l2_g40": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 2, "tile_gemm_ctas": 40}),
    "s_m64_l2_g56": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 2, "tile_gemm_ctas": 56}),
    "s_m64_l4_g12": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 4, "tile_gemm_ctas": 12}),
    "s_m64_l4_g24": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 4, "tile_gemm_ctas": 24}),
    "s_m64_l4_g40": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 4, "tile_gemm_ctas": 40}),
    "s_m64_l4_g56": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 4, "tile_gemm_ctas": 56}),
    "s_m64_l8_g12": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 8, "tile_gemm_ctas": 12}),
    "s_m64_l8_g24": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 8, "tile_gemm_ctas": 24}),
    "s_m64_l8_g40": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 8, "tile_gemm_ctas": 40}),
    "s_m64_l8_g56": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 8, "tile_gemm_ctas": 56}),
    "s_m128_l2_g12": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 2, "tile_gemm_ctas": 12}),
    "s_m128_l2_g24": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 2, "tile_gemm_ctas": 24}),
    "s_m128_l2_g40": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 2, "tile_gemm_ctas": 40}),
    "s_m128_l2_g56": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 2, "tile_gemm_ctas": 56}),
    "s_m128_l4_g12": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 4, "tile_gemm_ctas": 12}),
    "s_m128_l4_g24": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 4, "tile_gemm_ctas": 24}),
    "s_m128_l4_g40": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 4, "tile_gemm_ctas": 40}),
    "s_m128_l4_g56": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 4, "tile_gemm_ctas": 56}),
    "s_m128_l8_g12": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 8, "tile_gemm_ctas": 12}),
    "s_m128_l8_g24": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 8, "tile_gemm_ctas": 24}),
    "s_m128_l8_g40": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 8, "tile_gemm_ctas": 40}),
    "s_m128_l8_g56": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 8, "tile_gemm_ctas": 56}),
    "n_g8": (L.F_FORCE_TILE, {"tile_gemm_ctas": 8}),
    "n_g12": (L.F_FORCE_TILE, {"tile_gemm_ctas": 12}),
    "n_g16": (L.F_FORCE_TILE, {"tile_gemm_ctas": 16}),
    "n_g24": (L.F_FORCE_TILE, {"tile_gemm_ctas": 24}),
    "n_g32": (L.F_FORCE_TILE, {"tile_gemm_ctas": 32}),
    "n_g40": (L.F_FORCE_TILE, {"tile_gemm_ctas": 40}),
    "n_auto": (L.F_FORCE_TILE, {}),
    "n_ks8_g24": (L.F_FORCE_TILE, {"tile_gemm_ctas": 24, "tile_ksplit_tiles": 8}),
    "q_m12_l1": (L.F_FORCE_TILE, {"tile_m": 12, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m12_l2": (L.F_FORCE_TILE, {"tile_m": 12, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m12_l3": (L.F_FORCE_TILE, {"tile_m": 12, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "q_m16_l1": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m16_l2": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m16_l3": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "q_m20_l1": (L.F_FORCE_TILE, {"tile_m": 20, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m20_l2": (L.F_FORCE_TILE, {"tile_m": 20, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m20_l3": (L.F_FORCE_TILE, {"tile_m": 20, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "q_m24_l1": (L.F_FORCE_TILE, {"tile_m": 24, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m24_l2": (L.F_FORCE_TILE, {"tile_m": 24, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m24_l3": (L.F_FORCE_TILE, {"tile_m": 24, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "q_m28_l1": (L.F_FORCE_TILE, {"tile_m": 28, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m28_l2": (L.F_FORCE_TILE, {"tile_m": 28, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m28_l3": (L.F_FORCE_TILE, {"tile_m": 28, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "q_m32_l1": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m32_l2": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m32_l3": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "q_m40_l1": (L.F_FORCE_TILE, {"tile_m": 40, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m40_l2": (L.F_FORCE_TILE, {"tile_m": 40, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m40_l3": (L.F_FORCE_TILE, {"tile_m": 40, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "q_m48_l1": (L.F_FORCE_TILE, {"tile_m": 48, "tile_lag": 1, "tile_gemm_ctas": 6}),
    "q_m48_l2": (L.F_FORCE_TILE, {"tile_m": 48, "tile_lag": 2, "tile_gemm_ctas": 6}),
    "q_m48_l3": (L.F_FORCE_TILE, {"tile_m": 48, "tile_lag": 3, "tile_gemm_ctas": 6}),
    "tile_nostore": (L.F_FORCE_TILE, {"tile_nodeps": 4, "tile_gemm_ctas": 8}),
    "tile_noop": (L.F_FORCE_TILE, {"tile_nodeps": 2, "tile_gemm_ctas": 8}),
    "tile_noop_c14": (L.F_FORCE_TILE, {"tile_nodeps": 2, "tile_gemm_ctas": 8, "tile_chunk_kb": 14}),
    "tile_noop_c50": (L.F_FORCE_TILE, {"tile_nodeps": 2, "tile_gemm_ctas": 8, "tile_chunk_kb": 50}),
    "tile_nodeps_g1": (L.F_FORCE_TILE, {"tile_nodeps": 1, "tile_gemm_ctas": 1}),
    "t_m16_l4_g8": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 4, "tile_gemm_ctas": 8}),
    "t_m16_l6_g8": (L.F_FORCE_TILE, {"tile_m": 16, "tile_lag": 6, "tile_gemm_ctas": 8}),
    "t_m8_l8_g8": (L.F_FORCE_TILE, {"tile_m": 8, "tile_lag": 8, "tile_gemm_ctas": 8}),
    "t_m32_l3_g8": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 3, "tile_gemm_ctas": 8}),
    "t_m32_l4_g12": (L.F_FORCE_TILE, {"tile_m": 32, "tile_lag": 4, "tile_gemm_ctas": 12}),
    "t_m64_l3_g12": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 3, "tile_gemm_ctas": 12}),
    "t_m64_l4_g16": (L.F_FORCE_TILE, {"tile_m": 64, "tile_lag": 4, "tile_gemm_ctas": 16}),
    "t_m128_l3_g40": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 3, "tile_gemm_ctas": 40}),
    "t_m128_l4_g48": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 4, "tile_gemm_ctas": 48}),
    "t_m128_l2_g48": (L.F_FORCE_TILE, {"tile_m": 128, "tile_lag": 2, "tile_gemm_ctas": 48}),
    "tile_c50": (L.F_FORCE_TILE, {"tile_chunk_kb": 50}),
    "auto_biggemm": (0, {"gemm_big_tiles": 1}),
    "auto_ffma": (0, {"gemm_tf32x3": 0, "gemm_umma": 0}),
    "auto_mmasync": (0, {"gemm_umma": 0}),
    "stream_ffma": (L.F_FORCE_STREAMING, {"l2_chunk_mb": 100000, "gemm_tf32x3": 0, "gemm_umma": 0}),
    "stream_c40": (L.F_FORCE_STREAMING, {"l2_chunk_mb": 40}),
    "stream_c64": (L.F_FORCE_STREAMING, {"l2_chunk_mb": 64}),
    "stream_c100": (L.F_FORCE_STREAMING, {"l2_chunk_mb": 100}),
    "stream_nochunk": (L.F_FORCE_STREAMING, {"l2_chunk_mb": 100000}),
    "fused_cs4_t512": (L.F_FORCE_FUSED, {"fused_kind": 1, "fused_cluster": 4, "fused_threads": 512}),
    "fused_cs4_t256": (L.F_FORCE_FUSED, {"fused_kind": 1, "fused_cluster": 4, "fused_threads": 256}),
    "fused_cs8_t256": (L.F_FORCE_FUSED, {"fused_kind": 1, "fused_cluster": 8, "fused_threads": 256}),
    "l2_cs8": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 8}),
    "l2_cs8_occ5": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 8, "fused_occ": 5}),
    "l2_cs4_occ5": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 4, "fused_occ": 5}),
    "l2_cs4_anyw": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 4, "fused_weight_ratio_x100": 100000}),
    "l2_cs8_anyw": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 8, "fused_weight_ratio_x100": 100000}),
    "l2_cs8_g2": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 8, "fused_group_kb": 256}),
    "l2_cs4_g1": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 4, "fused_group_kb": 64}),
    "l2_cs8_g1": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 8, "fused_group_kb": 64, "fused_weight_ratio_x100": 100000}),
    "l2_pf37": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_prefetch": 37}),
    "l2_pf74": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_prefetch": 74}),
    "l2_pf148": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_prefetch": 148}),
    "l2_auto": (L.F_FORCE_FUSED, {"fused_kind": 2}),
    "l2_cs16": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 16}),
    "l2_cs16_g2": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 16, "fused_group_kb": 256}),
    "l2_stash0": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_stash_kb": 0}),
    "l2_stash24": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_stash_kb": 24}),
    "l2_stash46": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_stash_kb": 46}),
    "l2_stash100": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_stash_kb": 100}),
    "l2_cs8_nopf": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 8, "fused_prefetch": 0}),
    "l2_cs4_nopf": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 4, "fused_prefetch": 0}),
    "l2_cs4": (L.F_FORCE_FUSED, {"fused_kind": 2, "fused_cluster": 4}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="32,256,1024")
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--variants", default=",".join(VARIANTS))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--profile", action="store_true", help="print the library's per-kernel-class times per variant")
    ap.add_argument("--shapes", default="", help="comma-separated channel counts to keep (default: all three blocks)")
    args = ap.parse_args()
    keep = [int(x) for x in args.shapes.split(",") if x]
    lib = L.load()
    dev = torch.device("cuda:0")
    peak, _ = measured_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    rows = []
    for c, h in SHAPES:
        if keep and c not in keep:
            continue
        for n in [int(x) for x in args.batches.split(",")]:
            b = BlockBuffers(torch, L, n, c, h, dev, seed=c)
            u = n * c * h * h * 4
            P = lambda t: t.data_ptr()
            w, dw = b.w, b.dw

            def fwd(flags):
                return lib.gml_mmtm_fwd(P(b.a), P(b.b), P(b.a_out), P(b.b_out), P(w[0]), P(w[1]), P(w[2]), P(w[3]),
                                        P(w[4]), P(w[5]), P(b.z), P(b.hid), P(b.g_a), P(b.g_b), P(b.gate_sum),
                                        P(b.run_v), P(b.run_s), 0, None, None, P(b.fws), b.fws_bytes, b.dims, 0, 1.0,
                                        flags, stream.cuda_stream)

            def bwd(flags):
                return lib.gml_mmtm_bwd(P(b.go_a), P(b.go_b), P(b.a), P(b.b), P(w[0]), P(w[2]), P(w[4]), P(b.z),
                                        P(b.hid), P(b.g_a), P(b.g_b), None, None, None, None, P(b.d_a), P(b.d_b),
                                        P(dw[0]), P(dw[1]), P(dw[2]), P(dw[3]), P(dw[4]), P(dw[5]), P(b.ws),
                                        b.ws_bytes, b.dims, 0, 1.0, flags, stream.cuda_stream)

            for name in args.variants.split(","):
                flags, tun = VARIANTS[name]
                for k, v in {"l2_chunk_mb": 100000, "fused_kind": 0, "fused_cluster": 0, "fused_threads": 0, "fused_prefetch": 0, "fused_occ": 4, "fused_weight_ratio_x100": 100, "fused_group_kb": 128, "gemm_big_tiles": 0, "fused_stash_kb": -1, "gemm_tf32x3": 1, "gemm_umma": 1, "tile_kind": 0, "tile_lag": 0, "tile_ksplit_tiles": 0, "tile_m": 0, "tile_gemm_ctas": 0, "tile_chunk_kb": 28, "tile_nodeps": 0, "tile_switch": 1, "tile_rpol": 0, "tile_max_slots": 0, "tile_split_copies": 1, "tile_draw": 4, "tile_chunk_kb_fwd": 56, "tile_min_mb_light": 190, "tile_wgrad": 1, "fused_wsmem": -1, "tile_light_fwd": 0, "fused_hw_special": 1, **tun}.items():
                    L.check(lib.gml_set_tunable(k.encode(), v))
                # workspace sizes depend on the tile tunables: re-query for this variant
                b.ws_bytes = lib.gml_mmtm_bwd_workspace_bytes(b.dims)
                b.ws = torch.empty(b.ws_bytes, dtype=torch.uint8, device=dev)
                b.fws_bytes = lib.gml_mmtm_fwd_workspace_bytes(b.dims)
                b.fws = torch.empty(b.fws_bytes, dtype=torch.uint8, device=dev)
                rc = fwd(flags)
                if rc == -5:
                    continue  # unsupported shape for this variant
                L.check(rc, "fwd")
                L.check(bwd(flags), "bwd")
                torch.cuda.synchronize()
                res = {}
                for what, fn, units in (("fwd", fwd, 4), ("bwd", bwd, 6)):
                    ts = []
                    for _ in range(args.iters + 2):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        L.check(fn(flags), what)
                        e1.record()
                        e1.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ts = sorted(ts[2:])
                    med = ts[len(ts) // 2]
                    res[what] = {"ms": med, "gbs": units * u / (med * 1e-3) / 1e9, "min_ms": ts[0]}
                if args.profile:
                    import ctypes
                    for what, fn in (("fwd", fwd), ("bwd", bwd)):
                        lib.gml_profile_reset()
                        lib.gml_profile_enable(1)
                        for _ in range(4):
                            flush.zero_()
                            L.check(fn(flags), what)
                        torch.cuda.synchronize()
                        lib.gml_profile_enable(0)
                        parts = []
                        for tag in range(lib.gml_kernel_tag_count()):
                            t_, c_ = ctypes.c_double(), ctypes.c_int64()
                            lib.gml_profile_read(tag, ctypes.byref(t_), ctypes.byref(c_))
                            if c_.value:
                                parts.append("%s %.3f ms x%d" % (lib.gml_kernel_tag_name(tag).decode(), t_.value / 4, c_.value // 4))
                        print("    %s %s: %s" % (name, what, "; ".join(parts)), flush=True)
                tot = res["fwd"]["ms"] + res["bwd"]["ms"]
                row = {"c": c, "h": h, "n": n, "variant": name, "fwd_ms": res["fwd"]["ms"], "bwd_ms": res["bwd"]["ms"],
                       "fwd_gbs": res["fwd"]["gbs"], "bwd_gbs": res["bwd"]["gbs"],
                       "fwd_bwd_gbs": 10 * u / (tot * 1e-3) / 1e9, "frac_peak": 10 * u / (tot * 1e-3) / 1e9 / peak}
                rows.append(row)
                print("C=%3d H=%2d N=%4d %-16s fwd %8.3f ms %6.0f GB/s | bwd %8.3f ms %6.0f GB/s | fwd+bwd %6.0f GB/s "
                      "= %4.1f%% of %.0f" % (c, h, n, name, row["fwd_ms"], row["fwd_gbs"], row["bwd_ms"],
                                             row["bwd_gbs"], row["fwd_bwd_gbs"], 100 * row["frac_peak"], peak),
                      flush=True)
            del b
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
