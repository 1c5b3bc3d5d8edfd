/* Prints sizeof / offsetof of every struct of include/pilco_b200.h as the C compiler lays them out; the CPU test
 * compares the numbers with the ctypes mirrors in pilco_b200/_lib.py. */
#include <stdio.h>
#include <stddef.h>
#include "../../include/pilco_b200.h"
#define S(T) printf("sizeof %s %zu\n", #T, sizeof(T))
#define O(T, f) printf("offsetof %s.%s %zu\n", #T, #f, offsetof(T, f))
int main(void) {
    S(pilco_gp_model); O(pilco_gp_model, mode); O(pilco_gp_model, X); O(pilco_gp_model, X_bs); O(pilco_gp_model, beta);
    O(pilco_gp_model, iK); O(pilco_gp_model, ldk);
    S(pilco_policy); O(pilco_policy, squash); O(pilco_policy, max_action); O(pilco_policy, W); O(pilco_policy, b_bs);
    O(pilco_policy, rbf);
    S(pilco_reward_term); O(pilco_reward_term, channel); O(pilco_reward_term, coef); O(pilco_reward_term, W);
    O(pilco_reward_term, t);
    S(pilco_rollout); O(pilco_rollout, dyn); O(pilco_rollout, pol); O(pilco_rollout, n_rewards); O(pilco_rollout, rewards);
    O(pilco_rollout, m0); O(pilco_rollout, S0_bs); O(pilco_rollout, traj_m); O(pilco_rollout, reward);
    O(pilco_rollout, step_reward); O(pilco_rollout, info); O(pilco_rollout, ws); O(pilco_rollout, ws_bytes);
    O(pilco_rollout, mult_mu); O(pilco_rollout, step_risk); O(pilco_rollout, tape); O(pilco_rollout, tape_bytes);
    S(pilco_rollout_grad); O(pilco_rollout_grad, gb); O(pilco_rollout_grad, pol_L); O(pilco_rollout_grad, gm0);
    O(pilco_rollout_grad, ws); O(pilco_rollout_grad, ws_bytes);
    printf("version %d\n", PILCO_ABI_VERSION);
    return 0;
}
