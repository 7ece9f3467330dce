"""torch-CPU restatement of the reference's ForwardKL sampled-action path, executed the way the
reference executes it (materialised [B*N,S] / [B*N,A] stacks, nn.Linear layers, ~10 [B,N]
temporaries) -- the CPU arm that bench.py times beside the GPU numbers.

TEST INFRASTRUCTURE ONLY (see oracle/oracle_np.py): only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this.  /root/reference is not present
on the GPU box, so this port -- checked against the reference's own classes through the golden
vectors (tests/test_oracle.py::test_torch_port_matches_reference_golden) -- is what gets timed there
(cpu_baseline.kind == "port").
"""
from __future__ import annotations

import torch


class SoftQNetworkPort(torch.nn.Module):
    """``SoftQNetwork`` (forwardkl_network.py:250-268): cat([s,a],1) -> Linear+ReLU -> Linear+ReLU
    -> Linear; weights in torch [out,in] layout."""

    def __init__(self, W1, b1, W2, b2, W3, b3):
        super().__init__()
        t = lambda x: torch.nn.Parameter(torch.as_tensor(x, dtype=torch.float32).clone(), requires_grad=False)
        self.W1, self.b1, self.W2, self.b2, self.W3, self.b3 = t(W1), t(b1), t(W2), t(b2), t(W3), t(b3)

    def forward(self, state, action):
        x = torch.cat([state, action], 1)                                   # :264
        x = torch.relu(torch.nn.functional.linear(x, self.W1, self.b1))     # :265
        x = torch.relu(torch.nn.functional.linear(x, self.W2, self.b2))     # :266
        return torch.nn.functional.linear(x, self.W3.reshape(1, -1), self.b3.reshape(1))  # :267


def get_logprob_port(mean, log_std, grid_actions, action_scale, epsilon=1e-6):
    """``PolicyNetwork.get_logprob`` (forwardkl_network.py:324-351) given the head outputs
    mean/log_std [B,A] (the two-layer actor trunk is per-state work outside the path): tile the
    grid to [B,N,A], atanh, Normal / MultivariateNormal(mean, diag_embed(std)) log_prob, tanh
    Jacobian.  Returns [B,N]."""
    B, N = mean.shape[0], grid_actions.shape[0]
    tiled = grid_actions.unsqueeze(0).repeat(B, 1, 1)                                    # :104-106
    na = tiled.permute(1, 0, 2) / action_scale                                           # :329
    at = (torch.log(1 + na) - torch.log(1 - na)) / 2                                     # :330,353-354
    std = log_std.exp()                                                                  # :333
    if mean.shape[1] == 1:
        normal = torch.distributions.Normal(mean, std)                                   # :347-348
    else:
        normal = torch.distributions.MultivariateNormal(mean, torch.diag_embed(std))     # :350
    lp = normal.log_prob(at)                                                             # :337
    if lp.dim() == 2:
        lp = lp.unsqueeze(-1)
    lp = lp - torch.log(1 - na.pow(2) + epsilon).sum(dim=-1, keepdim=True)               # :342
    return lp.permute(1, 0, 2).reshape(B, N)                                             # :343


def fkl_sampled_step(qnet: SoftQNetworkPort, state, grid_actions, grid_weights, logp, entropy_scale,
                     action_scale=1.0):
    """The hot loop of ``ForwardKLNetwork.update_network`` (forwardkl_network.py:160-194):
    stack states, tile the grid, evaluate Q on B*N rows, Boltzmann-normalise per state, evaluate
    the policy log-density on the grid and form the per-state loss.  ``logp`` is either the
    [B,N] log-probabilities or the policy head outputs ``(mean, log_std)`` [B,A] each (then
    ``get_logprob_port`` runs inside, as in the reference).  Returns (loss_b [B], q [B,N])."""
    B, N = state.shape[0], grid_actions.shape[0]
    with torch.no_grad():
        if isinstance(logp, (tuple, list)):
            logp = get_logprob_port(logp[0], logp[1], grid_actions, action_scale)
        stacked_s = state.unsqueeze(1).repeat(1, N, 1).reshape(-1, state.shape[1])       # :161-162
        stacked_a = grid_actions.repeat(B, 1, 1).reshape(-1, grid_actions.shape[1])      # :104-105
        q = qnet(stacked_s, stacked_a).reshape(B, N)                                     # :164
        scaled = q / entropy_scale                                                       # :171
        m = scaled.max(dim=1, keepdim=True)[0]                                           # :173
        e = torch.exp(scaled - m)                                                        # :176
        z = (e * grid_weights).sum(dim=1, keepdim=True)                                  # :179-181
        boltz = e / z                                                                    # :184
        loss_b = -(grid_weights * boltz * logp).sum(dim=1)                               # :190-194
    return loss_b, q
