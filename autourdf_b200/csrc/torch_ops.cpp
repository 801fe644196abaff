// torch_ops.cpp -- the thin torch extension over the C ABI ("Torch extension" row).
//
// TORCH_LIBRARY(aurdf, ...) operators that check device / dtype / contiguity, allocate outputs and workspace as
// torch tensors, fetch at::cuda::getCurrentCUDAStream() and call libaurdf.so's extern "C" entry points -- nothing
// else.  Built by g++ (no nvcc, no Python.h) into autourdf_b200/libaurdf_torch.so and loaded with
// torch.ops.load_library (autourdf_b200/torch_ops.py).  Replaces, behind torch.ops.aurdf.*:
//   icp_sweep              masked_icp's sweep, AutoURDF PointCloud/cluster_icp.py:118-191 (device tensors)
//   se3_apply              calculate_pc, PointCloud/mlp_reg.py:155-170 (autograd: aurdf_se3_apply_bwd)
//   nn_l2                  open3d's correspondence search alone
//   dq_op                  the 11 functions of PointCloud/dq_func.py + the 4 pytorch3d functions (autograd:
//                          aurdf_dq_op_bwd); dualquat_to_transform / transform_to_dualquat / quaternion_to_matrix /
//                          matrix_to_quaternion are registered by name as well (the pairs train() differentiates
//                          through, mlp_reg.py:60-66, 78-84)
//   chamfer_distance       pytorch3d loss.chamfer_distance(x, y, norm) as mlp_reg.py:96 calls it (autograd)
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <string>
#include <tuple>
#include <vector>

#include "aurdf.h"

namespace {

using at::Tensor;

void check(int rc, const char *what) {
    TORCH_CHECK(rc == AURDF_OK, what, " failed with code ", rc, ": ", aurdf_last_error_string());
}
aurdf_stream_t stream_of(const Tensor &t) { return (aurdf_stream_t)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
void need_cuda(const Tensor &t, const char *name) { TORCH_CHECK(t.is_cuda() && t.is_contiguous(), name, ": contiguous CUDA tensor expected"); }
int pts_dtype(const Tensor &t) {
    TORCH_CHECK(t.scalar_type() == at::kFloat || t.scalar_type() == at::kDouble, "float32 / float64 points expected");
    return t.scalar_type() == at::kFloat ? AURDF_F32 : AURDF_F64;
}

// ---- cluster-ICP sweep --------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> icp_sweep(
    const Tensor &src, const Tensor &src_off, const Tensor &tgt, const Tensor &tgt_off, const Tensor &tile_frame,
    const c10::optional<Tensor> &box, const c10::optional<Tensor> &box_off, const Tensor &init_T, int64_t max_src_per_tile,
    int64_t tgt_capacity, double box_scale, double max_corr, int64_t max_iter, double rel_fitness, double rel_rmse,
    bool ori_only) {
    need_cuda(src, "src"); need_cuda(tgt, "tgt"); need_cuda(src_off, "src_off"); need_cuda(tgt_off, "tgt_off");
    need_cuda(tile_frame, "tile_frame"); need_cuda(init_T, "init_T");
    TORCH_CHECK(src.scalar_type() == tgt.scalar_type(), "src and tgt must share a dtype");
    TORCH_CHECK(src_off.scalar_type() == at::kInt && tgt_off.scalar_type() == at::kInt && tile_frame.scalar_type() == at::kInt,
                "offsets and tile_frame must be int32");
    TORCH_CHECK(init_T.scalar_type() == at::kDouble, "init_T must be float64 (B, 4, 4)");
    if (box.has_value()) {
        need_cuda(*box, "box");
        TORCH_CHECK(box_off.has_value() && box_off->scalar_type() == at::kInt, "box needs int32 box_off");
    }
    c10::cuda::CUDAGuard guard(src.device());
    const int64_t B = tile_frame.numel(), N = src.size(0);
    auto f64 = src.options().dtype(at::kDouble), i32 = src.options().dtype(at::kInt);
    Tensor T = at::empty({B, 4, 4}, f64), world = at::empty({N, 3}, f64), corr = at::empty({N}, i32);
    Tensor fit = at::empty({B}, f64), rmse = at::empty({B}, f64), iters = at::empty({B}, i32), ntgt = at::empty({B}, i32);
    Tensor status = at::zeros({4}, i32);
    if (B == 0) return {T, world, corr, fit, rmse, iters, ntgt, status};
    if (tgt_capacity <= 0) tgt_capacity = 4 * N + 2 * B + 1024;
    const size_t ws_bytes = aurdf_icp_workspace_bytes((int32_t)B, N, tgt_capacity);
    Tensor ws = at::empty({(int64_t)ws_bytes + 256}, src.options().dtype(at::kByte));
    char *wsp = (char *)ws.data_ptr();
    wsp += (256 - ((uintptr_t)wsp & 255)) & 255;
    check(aurdf_icp_sweep(src.data_ptr(), pts_dtype(src), src_off.data_ptr<int32_t>(), tgt.data_ptr(), tgt_off.data_ptr<int32_t>(),
                          tile_frame.data_ptr<int32_t>(), box.has_value() ? box->data_ptr() : nullptr,
                          box.has_value() ? pts_dtype(*box) : AURDF_F32, box.has_value() ? box_off->data_ptr<int32_t>() : nullptr,
                          init_T.data_ptr<double>(), (int32_t)B, N, (int32_t)max_src_per_tile, box_scale, max_corr, (int32_t)max_iter,
                          rel_fitness, rel_rmse, ori_only ? 1 : 0, T.data_ptr<double>(), world.data_ptr<double>(),
                          corr.data_ptr<int32_t>(), fit.data_ptr<double>(), rmse.data_ptr<double>(), iters.data_ptr<int32_t>(),
                          ntgt.data_ptr<int32_t>(), wsp, ws_bytes, tgt_capacity, status.data_ptr<int32_t>(), stream_of(src)),
          "aurdf_icp_sweep");
    return {T, world, corr, fit, rmse, iters, ntgt, status};
}

// ---- SE(3) apply (calculate_pc) -----------------------------------------------------------------------------
Tensor se3_apply_fwd(const Tensor &xyz, const Tensor &off, const Tensor &T) {
    need_cuda(xyz, "xyz"); need_cuda(off, "off"); need_cuda(T, "T");
    TORCH_CHECK(xyz.scalar_type() == T.scalar_type() && off.scalar_type() == at::kInt, "xyz / T share a dtype, off is int32");
    c10::cuda::CUDAGuard guard(xyz.device());
    Tensor out = at::empty_like(xyz);
    check(aurdf_se3_apply(xyz.data_ptr(), off.data_ptr<int32_t>(), T.data_ptr(), (int32_t)(off.numel() - 1), xyz.size(0), pts_dtype(xyz),
                          out.data_ptr(), stream_of(xyz)), "aurdf_se3_apply");
    return out;
}

struct Se3Apply : public torch::autograd::Function<Se3Apply> {
    static Tensor forward(torch::autograd::AutogradContext *ctx, const Tensor &xyz, const Tensor &off, const Tensor &T) {
        at::AutoDispatchBelowADInplaceOrView g;
        ctx->save_for_backward({xyz, off, T});
        return se3_apply_fwd(xyz.contiguous(), off, T.contiguous());
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext *ctx, torch::autograd::variable_list grads) {
        auto saved = ctx->get_saved_variables();
        Tensor xyz = saved[0].contiguous(), off = saved[1], T = saved[2].contiguous(), g = grads[0].contiguous();
        c10::cuda::CUDAGuard guard(xyz.device());
        Tensor gx = at::empty_like(xyz), gT = at::zeros_like(T);
        check(aurdf_se3_apply_bwd(g.data_ptr(), xyz.data_ptr(), off.data_ptr<int32_t>(), T.data_ptr(), (int32_t)(off.numel() - 1), xyz.size(0),
                                  pts_dtype(xyz), gx.data_ptr(), gT.data_ptr(), stream_of(xyz)), "aurdf_se3_apply_bwd");
        return {gx, Tensor(), gT};
    }
};
Tensor se3_apply_autograd(const Tensor &xyz, const Tensor &off, const Tensor &T) { return Se3Apply::apply(xyz, off, T); }

// ---- nearest neighbour (squared L2, float64 exact) ------------------------------------------------------------
std::tuple<Tensor, Tensor> nn_l2(const Tensor &query, const Tensor &q_off, const Tensor &target, const Tensor &t_off) {
    need_cuda(query, "query"); need_cuda(target, "target"); need_cuda(q_off, "q_off"); need_cuda(t_off, "t_off");
    TORCH_CHECK(query.scalar_type() == target.scalar_type(), "query and target must share a dtype");
    c10::cuda::CUDAGuard guard(query.device());
    Tensor idx = at::empty({query.size(0)}, query.options().dtype(at::kInt)), d2 = at::empty({query.size(0)}, query.options().dtype(at::kDouble));
    check(aurdf_nn_l2(query.data_ptr(), q_off.data_ptr<int32_t>(), target.data_ptr(), t_off.data_ptr<int32_t>(), pts_dtype(query),
                      (int32_t)(q_off.numel() - 1), query.size(0), idx.data_ptr<int32_t>(), d2.data_ptr<double>(), stream_of(query)), "aurdf_nn_l2");
    return {idx, d2};
}

// ---- dual-quaternion / quaternion operators ----------------------------------------------------------------
// op codes of aurdf_dq_op (include/aurdf.h); trailing shapes of (in0, in1, out0, out1); -1 = absent
struct DqShape { std::vector<int64_t> in0, in1, out0, out1; bool has_in1, has_out1; };
DqShape dq_shape(int64_t op) {
    switch (op) {
        case 0: return {{3, 3}, {3}, {4, 4}, {}, true, false};      // transform_from_rot_trans
        case 1: return {{4}, {}, {4}, {}, false, false};            // quaternion_conjugate
        case 2: return {{4}, {3}, {8}, {}, true, false};            // quat_trans_to_dualquat
        case 3: return {{3, 3}, {3}, {8}, {}, true, false};         // rot_trans_to_dualquat
        case 4: return {{4, 4}, {}, {8}, {}, false, false};         // transform_to_dualquat
        case 5: return {{8}, {}, {4}, {3}, false, true};            // dualquat_to_quat_trans
        case 6: return {{8}, {}, {3, 3}, {3}, false, true};         // dualquat_to_rot_trans
        case 7: return {{8}, {}, {4, 4}, {}, false, false};         // dualquat_to_transform
        case 8: return {{8}, {8}, {8}, {}, true, false};            // dualquat_multiply
        case 9: return {{8}, {}, {8}, {}, false, false};            // dualquat_invert
        case 10: return {{3}, {}, {8}, {}, false, false};           // point_to_dualquat
        case 11: return {{4}, {4}, {4}, {}, true, false};           // quaternion_raw_multiply
        case 12: return {{4}, {}, {4}, {}, false, false};           // quaternion_invert
        case 13: return {{4}, {}, {3, 3}, {}, false, false};        // quaternion_to_matrix
        case 14: return {{3, 3}, {}, {4}, {}, false, false};        // matrix_to_quaternion
    }
    TORCH_CHECK(false, "aurdf::dq_op: unknown operator ", op);
}
int64_t numel_of(const std::vector<int64_t> &s) { int64_t n = 1; for (auto v : s) n *= v; return n; }

std::tuple<Tensor, Tensor> dq_op_fwd(int64_t op, const Tensor &in0, const c10::optional<Tensor> &in1) {
    const DqShape sh = dq_shape(op);
    need_cuda(in0, "in0");
    TORCH_CHECK(sh.has_in1 == in1.has_value(), "aurdf::dq_op: operator ", op, sh.has_in1 ? " needs" : " takes no", " second operand");
    const int64_t n = in0.numel() / numel_of(sh.in0);
    TORCH_CHECK(n * numel_of(sh.in0) == in0.numel(), "aurdf::dq_op: in0 has the wrong trailing shape");
    if (in1.has_value()) {
        need_cuda(*in1, "in1");
        TORCH_CHECK(in1->numel() == n * numel_of(sh.in1) && in1->scalar_type() == in0.scalar_type(), "aurdf::dq_op: operands must be broadcast to one batch and share a dtype");
    }
    c10::cuda::CUDAGuard guard(in0.device());
    std::vector<int64_t> batch(in0.sizes().begin(), in0.sizes().end() - (int64_t)sh.in0.size());
    auto shape = [&](const std::vector<int64_t> &tail) { std::vector<int64_t> s = batch; s.insert(s.end(), tail.begin(), tail.end()); return s; };
    Tensor out0 = at::empty(shape(sh.out0), in0.options()), out1 = sh.has_out1 ? at::empty(shape(sh.out1), in0.options()) : at::empty({0}, in0.options());
    check(aurdf_dq_op((int)op, in0.data_ptr(), in1.has_value() ? in1->data_ptr() : nullptr, out0.data_ptr(), sh.has_out1 ? out1.data_ptr() : nullptr,
                      n, pts_dtype(in0), stream_of(in0)), "aurdf_dq_op");
    return {out0, out1};
}

struct DqOp : public torch::autograd::Function<DqOp> {
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext *ctx, int64_t op, const Tensor &in0, const c10::optional<Tensor> &in1) {
        at::AutoDispatchBelowADInplaceOrView g;
        Tensor a = in0.contiguous(), b = in1.has_value() ? in1->contiguous() : Tensor();
        ctx->saved_data["op"] = op;
        ctx->saved_data["has1"] = in1.has_value();
        ctx->save_for_backward({a, b});
        auto r = dq_op_fwd(op, a, in1.has_value() ? c10::optional<Tensor>(b) : c10::nullopt);
        return {std::get<0>(r), std::get<1>(r)};
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext *ctx, torch::autograd::variable_list grads) {
        auto saved = ctx->get_saved_variables();
        const int64_t op = ctx->saved_data["op"].toInt();
        const bool has1 = ctx->saved_data["has1"].toBool();
        const DqShape sh = dq_shape(op);
        Tensor in0 = saved[0], in1 = saved[1];
        c10::cuda::CUDAGuard guard(in0.device());
        const int64_t n = in0.numel() / numel_of(sh.in0);
        Tensor g0 = grads[0].defined() ? grads[0].contiguous() : Tensor(), g1 = (sh.has_out1 && grads[1].defined()) ? grads[1].contiguous() : Tensor();
        Tensor gin0 = at::empty_like(in0), gin1 = has1 ? at::empty_like(in1) : Tensor();
        check(aurdf_dq_op_bwd((int)op, in0.data_ptr(), has1 ? in1.data_ptr() : nullptr, g0.defined() ? g0.data_ptr() : nullptr,
                              g1.defined() ? g1.data_ptr() : nullptr, gin0.data_ptr(), has1 ? gin1.data_ptr() : nullptr, n, pts_dtype(in0),
                              stream_of(in0)), "aurdf_dq_op_bwd");
        return {Tensor(), gin0, gin1};
    }
};
std::tuple<Tensor, Tensor> dq_op_autograd(int64_t op, const Tensor &in0, const c10::optional<Tensor> &in1) {
    auto r = DqOp::apply(op, in0, in1);
    return {r[0], r[1]};
}
Tensor transform_to_dualquat(const Tensor &T) { return DqOp::apply(4, T, c10::nullopt)[0]; }
Tensor dualquat_to_transform(const Tensor &dq) { return DqOp::apply(7, dq, c10::nullopt)[0]; }
Tensor quaternion_to_matrix(const Tensor &q) { return DqOp::apply(13, q, c10::nullopt)[0]; }
Tensor matrix_to_quaternion(const Tensor &m) { return DqOp::apply(14, m, c10::nullopt)[0]; }

// ---- chamfer distance ------------------------------------------------------------------------------------
struct Chamfer : public torch::autograd::Function<Chamfer> {
    static Tensor forward(torch::autograd::AutogradContext *ctx, const Tensor &x_, const Tensor &y_, int64_t norm, bool point_mean, bool batch_mean) {
        at::AutoDispatchBelowADInplaceOrView g;
        Tensor x = x_.to(at::kFloat).contiguous(), y = y_.to(at::kFloat).contiguous();
        need_cuda(x, "x"); need_cuda(y, "y");
        TORCH_CHECK(x.dim() == 3 && y.dim() == 3 && x.size(0) == y.size(0) && x.size(2) == 3 && y.size(2) == 3, "x (N,P1,3), y (N,P2,3) expected");
        c10::cuda::CUDAGuard guard(x.device());
        const int32_t N = (int32_t)x.size(0), P1 = (int32_t)x.size(1), P2 = (int32_t)y.size(1);
        // a fresh, zeroed workspace per call keeps the operator free of hidden state (the Python wrapper caches one)
        const size_t wb = aurdf_chamfer_workspace_bytes(N, P1, P2);
        Tensor ws = at::zeros({(int64_t)wb + 256}, x.options().dtype(at::kByte));
        char *wsp = (char *)ws.data_ptr();
        wsp += (256 - ((uintptr_t)wsp & 255)) & 255;
        Tensor idx = at::empty({(int64_t)N * (P1 + P2)}, x.options().dtype(at::kInt)), loss = at::empty({}, x.options());
        check(aurdf_chamfer_fwd(x.data_ptr<float>(), y.data_ptr<float>(), N, P1, P2, (int)norm, point_mean, batch_mean, idx.data_ptr<int32_t>(),
                                idx.data_ptr<int32_t>() + (int64_t)N * P1, loss.data_ptr<float>(), wsp, wb, stream_of(x)), "aurdf_chamfer_fwd");
        ctx->save_for_backward({x, y, idx});
        ctx->saved_data["norm"] = norm;
        ctx->saved_data["pm"] = point_mean;
        ctx->saved_data["bm"] = batch_mean;
        return loss;
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext *ctx, torch::autograd::variable_list grads) {
        auto saved = ctx->get_saved_variables();
        Tensor x = saved[0], y = saved[1], idx = saved[2], gl = grads[0].to(at::kFloat).contiguous();
        c10::cuda::CUDAGuard guard(x.device());
        const int32_t N = (int32_t)x.size(0), P1 = (int32_t)x.size(1), P2 = (int32_t)y.size(1);
        Tensor gx = at::zeros_like(x), gy = at::zeros_like(y);
        check(aurdf_chamfer_bwd(x.data_ptr<float>(), y.data_ptr<float>(), idx.data_ptr<int32_t>(), idx.data_ptr<int32_t>() + (int64_t)N * P1,
                                gl.data_ptr<float>(), N, P1, P2, (int)ctx->saved_data["norm"].toInt(), ctx->saved_data["pm"].toBool(),
                                ctx->saved_data["bm"].toBool(), gx.data_ptr<float>(), gy.data_ptr<float>(), stream_of(x)), "aurdf_chamfer_bwd");
        return {gx, gy, Tensor(), Tensor(), Tensor()};
    }
};
Tensor chamfer_distance(const Tensor &x, const Tensor &y, int64_t norm, bool point_mean, bool batch_mean) {
    TORCH_CHECK(norm == 1 || norm == 2, "Support for 1 or 2 norm.");
    return Chamfer::apply(x, y, norm, point_mean, batch_mean);
}

}  // namespace

TORCH_LIBRARY(aurdf, m) {
    m.def("icp_sweep(Tensor src, Tensor src_off, Tensor tgt, Tensor tgt_off, Tensor tile_frame, Tensor? box, Tensor? box_off, Tensor init_T, "
          "int max_src_per_tile=0, int tgt_capacity=0, float box_scale=1.2, float max_corr=1.0, int max_iter=10000, float rel_fitness=1e-6, "
          "float rel_rmse=1e-6, bool ori_only=False) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)");
    m.def("se3_apply(Tensor xyz, Tensor off, Tensor T) -> Tensor");
    m.def("nn_l2(Tensor query, Tensor q_off, Tensor target, Tensor t_off) -> (Tensor, Tensor)");
    m.def("dq_op(int op, Tensor in0, Tensor? in1=None) -> (Tensor, Tensor)");   // second output: empty (0 elements) for single-output operators
    m.def("transform_to_dualquat(Tensor T) -> Tensor");
    m.def("dualquat_to_transform(Tensor dq) -> Tensor");
    m.def("quaternion_to_matrix(Tensor q) -> Tensor");
    m.def("matrix_to_quaternion(Tensor m) -> Tensor");
    m.def("chamfer_distance(Tensor x, Tensor y, int norm=2, bool point_mean=True, bool batch_mean=True) -> Tensor");
}

TORCH_LIBRARY_IMPL(aurdf, CUDA, m) {
    m.impl("icp_sweep", icp_sweep);
    m.impl("nn_l2", nn_l2);
}

// operators with an autograd formula: the Autograd key runs the Function (whose forward calls the C ABI below it)
TORCH_LIBRARY_IMPL(aurdf, Autograd, m) {
    m.impl("se3_apply", se3_apply_autograd);
    m.impl("dq_op", dq_op_autograd);
    m.impl("transform_to_dualquat", transform_to_dualquat);
    m.impl("dualquat_to_transform", dualquat_to_transform);
    m.impl("quaternion_to_matrix", quaternion_to_matrix);
    m.impl("matrix_to_quaternion", matrix_to_quaternion);
    m.impl("chamfer_distance", chamfer_distance);
}
