// nmpc_sim.cuh -- device memory map of the on-device closed loop (planner, plant, state estimation, logs).
#pragma once
#include "nmpc_kernels.cuh"

namespace b200 {

struct SimMem {
    // reference trajectory (Trajectories/reftraj_*.json: pos_x, pos_y, ref_yaw, ref_v) and the per-point travel
    // time to the successor |p[i+1] - p[i]| / v[i+1] (Utils/MPC_sim_utils.py:170)
    const double *ref_px, *ref_py, *ref_yaw, *ref_v, *seg_dt;
    int n_ref, maxw;            // maxw: longest time-horizon window over all start indices (shared-memory sizing)
    double Ts, Tp, jump_thresh; // jump_thresh = deg2rad(250)
    int nsub, cap;              // plant finite elements per Ts; log capacity in steps
    int win[8];                 // moving-average windows of StateEstimation
    // per-instance loop state
    double *pose;               // [B][4]  x, y, yaw, |v| of the undisturbed plant state
    double *cilx;               // [B][7]  plant state
    double *pred_x1;            // [B][8]  x_1 of the last successful solve
    double *fbuf;               // [B][8][4] filter ring buffers
    int *fcount;                // [B]
    int *ref_idx;               // [B] closest reference index of the last planner call
    const double *dist_deriv;   // [B][7] additive state-derivative disturbance of the next step (or null)
    const double *dist_est;     // [B][7] additive state-estimation noise of the next step (or null)
    // generated / played-back disturbances (Utils/MPC_sim_utils.py:16-99, SimulationMode_main_class.py:121-143)
    int gen_type[2];            // per channel (0 derivatives, 1 state estimation): 0 off, 1 uniform-in-ellipsoid,
                                // 2 gaussian, 3 absolute, 4 playback
    double gen_bound[2][7];     // upper disturbance bounds w_* of sim_main_params.yaml
    unsigned long long seed;    // counter-based generator key (Philox4x32-10 keyed by seed, counter = instance, step, channel)
    const double *play[2];      // [B][play_n][7] playback rows (type 4)
    int play_n;
    double *log_dist[2];        // [B][cap][7] sim_disturbance_derivatives / sim_disturbance_state_estimation
    // logs (Logging_Plotting.py:34-41), instance-major
    double *log_mpcx;           // [B][cap+1][8]
    double *log_cilx;           // [B][cap+1][7]  undisturbed plant state (Logger.CiLX)
    double *log_distx;          // [B][cap+1][7]  disturbed plant state + estimation noise input (Logger.DisturbedX)
    double *log_u;              // [B][cap][2]
    double *log_ref;            // [B][cap][4]
    double *log_dbg;            // [B][cap][5]
    int *log_idx;               // [B][cap] closest reference index of every step (current_ref_idx of PlannerEmulator)
    // stop conditions of the BO objective / RL environment (objective_function.py:153,186-200), evaluated per step
    int *stop_step;             // [B] -1 while running, else the number of steps simulated when the instance stopped
    int *stop_flags;            // [B] bit 0: closest index == end_idx (done), bit 1: crash predicate
    const int *end_idx;         // [B] segment end index (or null: no end condition)
    int stop_on;                // predicates enabled
    double max_lat_dev, max_a_comb;
};

int planner_smem_bytes(int maxw);
void launch_planner(const SimMem &s, double *yref, int B, int N, int step, cudaStream_t st);
void launch_sim_advance(const DevCfg &cfg, const DevMem &m, const SimMem &s, int step, cudaStream_t st);
void launch_sim_reset(const SimMem &s, const double *x0_mpc_dev, int B, cudaStream_t st);
void launch_snmpc_reinit_failed(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
void launch_sim_reset_masked(const DevCfg &cfg, const DevMem &m, const SimMem &s, const double *x0_mpc_dev, const int *mask_dev,
                             int step, cudaStream_t st);
void launch_launch_order(const DevMem &m, int *order, int B, cudaStream_t st);
void launch_snmpc_init_samples_masked(const DevCfg &cfg, const DevMem &m, const int *mask_dev, cudaStream_t st);

}  // namespace b200
