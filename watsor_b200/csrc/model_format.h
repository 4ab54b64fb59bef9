// model_format.h -- on-disk / in-memory layout of a compiled `.wb200` model blob.
// Written by watsor_b200/model.py (Model.to_blob), read by wb_create().
#pragma once
#include <stdint.h>

#define WB_MODEL_MAGIC "WB200M01"

enum wb_op : uint32_t {
  WB_OP_STEM = 1,
  WB_OP_DW = 2,
  WB_OP_PW = 3,
  WB_OP_CONV = 4,
  WB_OP_ADD = 5,
  WB_OP_HEAD = 6,
  WB_OP_MAXPOOL = 7,  // TF MaxPool, padding SAME
  WB_OP_AVGPOOL = 8,  // TF AvgPool, padding SAME (divides by the number of in-image taps)
  WB_OP_COPY = 9      // ConcatV2 along channels: in_c channels -> channels [row_off, row_off + in_c) of an out_c-wide tensor
};
enum wb_act : uint32_t { WB_ACT_NONE = 0, WB_ACT_RELU6 = 1 };

struct wb_model_header {  // 256 bytes
  char magic[8];
  uint32_t n_layers, n_tensors, input_h, input_w, num_classes, num_anchors;
  float pre_mul, pre_sub, scale_y, scale_x, scale_h, scale_w, logit_scale, iou_thr, score_thr;
  uint32_t max_per_class, max_total;
  float class_offset;
  uint32_t anchors_tensor, reserved0;
  uint64_t arena_elems;  // activation arena, elements per frame
  uint8_t pad[160];
};
static_assert(sizeof(wb_model_header) == 256, "header layout");

struct wb_layer {  // 128 bytes
  uint32_t op, act;
  uint32_t in_h, in_w, in_c, out_h, out_w, out_c;
  uint32_t kh, kw, stride, pad_t, pad_l;
  uint32_t in_off, in2_off, out_off;  // arena offsets, elements per frame
  int32_t w_tensor, scale_tensor, offset_tensor;
  uint32_t n_pad;                                   // leading dimension of the weight matrix
  uint32_t anchors_per_loc, row_off, n_box, n_cls;  // head layers
  char name[32];
};
static_assert(sizeof(wb_layer) == 128, "layer layout");

struct wb_tensor_entry {  // 16 bytes
  uint64_t offset;  // in floats from the start of the data section
  uint64_t count;
};
