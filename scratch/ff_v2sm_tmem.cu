#include <cuda_runtime.h>
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/util/packed_stride.hpp"

using namespace cute;

namespace v2sm_tmem {
using ElementA = float; using LayoutA = cutlass::layout::RowMajor;   constexpr int AlignA = 4;
using ElementB = float; using LayoutB = cutlass::layout::ColumnMajor; constexpr int AlignB = 4;
using ElementC = float; using LayoutC = cutlass::layout::RowMajor;   constexpr int AlignC = 4;
using ElementAcc = float;
using MmaTileShape = Shape<_256, _128, _16>;
using ClusterShape = Shape<_2, _1, _1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
    cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, AlignC,
    ElementC, LayoutC, AlignC, cutlass::epilogue::TmaWarpSpecialized2Sm>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, AlignA, ElementB, LayoutB,
    AlignB, ElementAcc, MmaTileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(
        sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32Sm100>::CollectiveOp;
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop,
                                                        CollectiveEpilogue>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
int run(int M, int N, int K, const float *A, const float *W, float *C, void *workspace,
        size_t workspace_bytes, cudaStream_t stream) {
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, 1));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, 1));
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1},
                                {A, sa, W, sb}, {{1.f, 0.f}, C, sc, C, sd}};
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return -1;
  if (Gemm::get_workspace_size(args) > workspace_bytes) return -2;
  if (gemm.initialize(args, workspace, stream) != cutlass::Status::kSuccess) return -3;
  if (gemm.run(stream) != cutlass::Status::kSuccess) return -4;
  return 0;
}
}

extern "C" int v2sm_tmem_gemm(int M, int N, int K, const float *A, const float *W, float *C, void *ws, size_t wsb, cudaStream_t st) { return v2sm_tmem::run(M, N, K, A, W, C, ws, wsb, st); }
