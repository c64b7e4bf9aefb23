"""Batched AG-DQN (DDQN with A*-guided exploration) driving thousands of scenes at once.

Reference: src/algorithm/MADQN_structure/MADQN.py:13-200 and Controller.py:21-145 (one AGV
decision, one stored transition and one optimiser step at a time).  The arithmetic of the
agent is unchanged -- same 6-conv CNN on the 3x7x7 limited-visual state (Net, MADQN.py:13-51),
epsilon = probability of using the network, otherwise the A* guidance action
(choose_action, 96-111), priority (|TD error| + 0.01)^0.6 set at insert (store_transition,
113-132; PER.py:79-80), DDQN target + SmoothL1 + Adagrad(lr 0.01), target sync every 100
updates (update_network, 134-171) -- but every tensor has a leading scene dimension and the
environment side is the CUDA path:

    agv_begin_tick
    for k in AGVs:   obs  = agv_encode(k)                      # create_state, bf16 [E,3,7,7]
                     a    = u < eps ? argmax Q(obs) : -1       # -1 = "A* guidance" ...
                     r, done = agv_step_turn(k, a)             # ... resolved inside the kernel
                     obs' = agv_encode(k, post=1)
                     agv_replay_push(obs, a, r, obs', done, priority)
                     learner step(s) on agv_replay_sample
    agv_end_tick

The CNN is the only dense contraction and stays in PyTorch (bf16 autocast, tensor cores); the
learner step is captured into CUDA graphs (two graphs around the gradient all-reduce with several ranks).
Multi-GPU: scenes and replay shard per rank; the flattened gradient is all-reduced (mean) over NCCL
once per update.  Known reference defect kept out of the data: a guidance STOP (action 4) cannot
index the 4-wide Q row (IndexError in MADQN.py:118), so STOP transitions are not stored.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib as L
from ...replay import DeviceReplay
from ...utils.utils import Direction
from .PER import Memory


class Net(nn.Module):
    """MADQN.py:13-51: conv 3->64->64->128->128->256->256 (3x3, pad 1, ReLU), one 2x2 max-pool
    (only conv_pool3 is applied in forward), fc 2304->256->actions."""

    def __init__(self, num_inputs=3, num_actions=4, map_xdim=7, map_ydim=7):
        super().__init__()
        self.conv1 = nn.Conv2d(num_inputs, 64, 3, 1, 1)
        self.conv2 = nn.Conv2d(64, 64, 3, 1, 1)
        self.conv_pool1 = nn.MaxPool2d(2, 2)
        self.conv3 = nn.Conv2d(64, 128, 3, 1, 1)
        self.conv4 = nn.Conv2d(128, 128, 3, 1, 1)
        self.conv_pool2 = nn.MaxPool2d(2, 2)
        self.conv5 = nn.Conv2d(128, 256, 3, 1, 1)
        self.conv6 = nn.Conv2d(256, 256, 3, 1, 1)
        self.conv_pool3 = nn.MaxPool2d(2, 2)
        self.fc3 = nn.Linear(2304, 256)
        self.fc4 = nn.Linear(256, num_actions)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        x = F.relu(self.conv4(x))
        x = F.relu(self.conv5(x))
        x = F.relu(self.conv6(x))
        x = self.conv_pool3(x)
        x = F.relu(self.fc3(x.reshape(x.size(0), -1)))
        return self.fc4(x)


def ddqn_loss(policy, target, obs, act, rew, nxt, done, gamma, loss_fn):
    """MADQN.py:150-164: Q(s,a) against r + gamma * Q_target(s', argmax_a' Q_policy(s', a')) * (1 - done).
    The bootstrap side carries no gradient to the policy parameters in the reference either (the argmax is an
    index, the target network is not in the optimiser), so it runs under no_grad."""
    q_sa = policy(obs).gather(1, act)
    with torch.no_grad():
        greedy = policy(nxt).argmax(1, keepdim=True)
        y = target(nxt).gather(1, greedy) * gamma * (1 - done) + rew
    return loss_fn(q_sa, y)


class Agent:
    """The reference's one-decision-at-a-time agent (MADQN.py:54-200) over this package's pieces: the
    prioritised replay is the device-resident Memory mirror (PER.py), the guidance action comes from the
    CUDA BFS behind the FindPathAstar mirror.  Same hyper-parameters, same call order on `np.random`
    (epsilon test) and `random` (PER sampling), torch.optim.Adagrad -- it is what MADQNAgentController
    drives through the Scene callback protocol, and the yardstick the batched learner is tested against.
    Reference defect not reproduced: a guidance STOP (action 4) indexes the 4-wide Q row there
    (IndexError, MADQN.py:118); here such a transition is not stored."""

    def __init__(self, policy_net, target_net, device):
        self.device = device
        self.policy_network = policy_net.to(self.device)
        self.target_network = target_net.to(self.device)
        self.epsilon_start = self.epsilon = 0.8
        self.epsilon_end, self.epsilon_count = 1, 0
        self.memory_size = 50000
        self.start_training_info_number = 100
        self.learn_step_counter = 0
        self.TARGET_REPLACE_ITER = 100
        self.batch_size = 32
        self.GAMMA = 0.95
        dev_index = self.device.index if getattr(self.device, "type", "cpu") == "cuda" else None
        self.memory = Memory(self.memory_size, device=dev_index)
        self.lr_start = self.lr = 0.01
        self.lr_end, self.lr_count = 0.001, 0
        self.optim = torch.optim.Adagrad(self.policy_network.parameters(), self.lr)
        self.loss_function = torch.nn.SmoothL1Loss()
        self.loss_value = []
        self.dir = Direction()

    def _q(self, net, obs):
        state = torch.from_numpy(np.asarray(obs)).float().unsqueeze(0).to(self.device)
        return net(state)

    def choose_action(self, obs, current_place, target_place, valid_path_matrix, epsilon=0.):
        epsilon = max(epsilon, self.epsilon)
        if np.random.uniform() < epsilon:
            return torch.max(self._q(self.policy_network, obs).cpu(), 1)[1].data.numpy()
        return np.array([self.find_action_astar(valid_path_matrix, current_place, target_place)])

    def store_transition(self, s, a, r, s_, is_done):
        a = int(a)
        q = self._q(self.policy_network, s).cpu()
        if a >= q.shape[1]:
            return  # guidance STOP: no Q entry to compare with (see the class docstring)
        old_val = q[0][a]
        target_val = self._q(self.policy_network, s_)
        new_val = r if is_done == 1 else r + self.GAMMA * torch.max(target_val)
        error = abs(old_val - new_val)
        error = error.detach().cpu().numpy() if torch.is_tensor(error) else np.asarray(error)
        self.memory.add(error, (np.array(s), a, r, np.array(s_), is_done))
        if self.memory.tree.n_entries >= self.start_training_info_number:
            self.update_network()

    def update_network(self):
        if self.learn_step_counter % self.TARGET_REPLACE_ITER == 0:
            self.target_network.load_state_dict(self.policy_network.state_dict())
        self.learn_step_counter += 1
        b = self.memory.sample_device(self.batch_size)  # Memory.sample: one random.random() per segment
        to = lambda t: t.to(self.device)
        loss = ddqn_loss(self.policy_network, self.target_network, to(b["obs"].float()),
                         to(b["act"].long().unsqueeze(1)), to(b["reward"].unsqueeze(1)), to(b["next_obs"].float()),
                         to(b["done"].long().unsqueeze(1)), self.GAMMA, self.loss_function)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        self.loss_value.append(min(0.5, max(-0.5, loss.item())))

    def change_learning_rate(self, times):
        """MADQN.py:173-180 (like the reference this only moves the Python float, not the optimiser)"""
        if self.lr_count > times:
            return
        self.lr = self.lr - (self.lr_start - self.lr_end) / times
        self.lr_count += 1

    def change_explore_rate(self, times):
        if self.epsilon_count >= times:
            self.epsilon = self.epsilon_end
        else:
            self.epsilon = self.epsilon + (self.epsilon_end - self.epsilon_start) / times
        self.epsilon_count += 1

    def find_action_astar(self, matrix_valid_map, current_position, target_position):
        from ...utils.astar import FindPathAstar
        finder = FindPathAstar(matrix_valid_map, (current_position[0] - 1, current_position[1] - 1),
                               (target_position[0] - 1, target_position[1] - 1))
        found, _path, _map, actions = finder.run_astar_method()
        return self.dir.action_str_value("STOP" if (found is False or len(actions) == 0) else actions[0])


class BatchedAgent:
    def __init__(self, scene, policy_net=None, target_net=None, memory_size=1 << 20, batch_size=256,
                 start_training_info_number=100, gamma=0.95, lr=0.01, target_replace_iter=100, epsilon=0.8,
                 updates_per_substep=1, autocast=True, ddp=False, seed=0, graph_learner=None,
                 updates_per_transition=None):
        self.scene = scene
        self.device = scene.device
        self.policy_network = (policy_net or Net()).to(self.device).to(memory_format=torch.channels_last)
        self.target_network = (target_net or Net()).to(self.device).to(memory_format=torch.channels_last)
        self.target_network.load_state_dict(self.policy_network.state_dict())
        self._train_net = self.policy_network
        self.world, self.rank = 1, 0
        if ddp:
            # data-parallel learner without the DDP wrapper (its hooks cannot be captured): same initial
            # weights on every rank, one all-reduce of the flattened gradient per update, mean over ranks
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
            for q in self.policy_network.parameters():
                dist.broadcast(q.data, src=0)
            self.target_network.load_state_dict(self.policy_network.state_dict())
        self.epsilon_start = self.epsilon = epsilon
        self.epsilon_end, self.epsilon_count = 1.0, 0
        self.memory_size, self.batch_size = memory_size, batch_size
        self.start_training_info_number = start_training_info_number
        self.learn_step_counter = 0
        self.TARGET_REPLACE_ITER = target_replace_iter
        self.GAMMA = gamma
        # learner schedule: a fixed number of updates per sub-step (default: 1, each on a batch drawn from all the
        # scenes' transitions), or -- the reference's schedule, MADQN.py:131-132 -- `updates_per_transition`
        # updates for every transition the sub-step stored (one synchronising count read per sub-step)
        self.updates_per_substep = updates_per_substep
        self.updates_per_transition = updates_per_transition
        self._upd_carry = 0.0
        self.memory = DeviceReplay(memory_size, (3, scene.K, scene.K), device=scene.device_index)
        # Adagrad(lr) with torch's defaults (lr_decay 0, eps 1e-10, accumulator 0), written as the same
        # three foreach ops torch.optim.Adagrad runs, so that the whole update can sit in a CUDA graph
        self.lr, self.adagrad_eps = lr, 1e-10
        self._params = [p for p in self.policy_network.parameters()]
        self._sums = [torch.zeros_like(p) for p in self._params]
        self.loss_function = nn.SmoothL1Loss()
        # the learner step (PER sample -> DDQN target -> SmoothL1 -> backward -> Adagrad) is ~150 tiny
        # kernels: replayed from CUDA graphs (2.5 ms of launch overhead -> ~1 ms of GPU work).  With
        # several ranks it is two graphs around one NCCL all-reduce of the flat gradient.
        self.graph_learner = True if graph_learner is None else bool(graph_learner)
        self._graph = None
        self._graph_b = None
        self._flat = torch.zeros(sum(q.numel() for q in self._params), device=self.device)
        self._views, off = [], 0
        for q in self._params:
            self._views.append(self._flat[off:off + q.numel()].view_as(q))
            off += q.numel()
        self._warm = 0
        self._u = torch.zeros(batch_size, dtype=torch.float64, device=self.device)
        self._sb = self.memory._alloc_batch(batch_size)
        self._loss = torch.zeros((), device=self.device)
        self.autocast = autocast
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed + self.rank)  # every rank draws its own epsilon tests / PER uniforms
        # losses of the last LOSS_RING updates stay on the device (no allocation, no sync per update)
        self._loss_ring = torch.zeros(self.LOSS_RING, device=self.device)
        self._stored_dev = torch.zeros((), dtype=torch.int64, device=self.device)  # transitions accepted by the replay
        self._learning = False  # the replay holds >= start_training_info_number entries (MADQN.py:131)
        # several ranks: the all-reduce of update u runs on a side stream while the main stream goes on with
        # the next sub-step; the Adagrad half of update u is applied right before update u+1 starts
        self._comm = torch.cuda.Stream(self.device) if self.world > 1 else None
        self._pending = None
        self.overlap_allreduce = True
        self.reward_acc = torch.zeros((), dtype=torch.float64, device=self.device)

    LOSS_RING = 4096

    @property
    def loss_value(self):
        """losses of the most recent updates (oldest first), clamped to +-0.5 like MADQN.py:165-171"""
        n = min(self.learn_step_counter, self.LOSS_RING)
        if n == 0:
            return []
        idx = [(self.learn_step_counter - n + j) % self.LOSS_RING for j in range(n)]
        return self._loss_ring[idx].clamp(-0.5, 0.5).cpu().tolist()

    def transitions_stored(self) -> int:
        """transitions this rank has put into its replay shard (synchronises)"""
        return int(self._stored_dev.item())

    # ------------------------------------------------------------------ acting
    def q_values(self, net, obs):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            x = obs if self.autocast else obs.float()
            return net(x.contiguous(memory_format=torch.channels_last)).float()

    def tick(self, train=True, epsilon=None, u=None):
        """one Scene.run_mode pass of every scene with the network in the loop; returns the
        number of AGV turns taken.  `u` (optional [N, E] uniforms) injects the epsilon draws."""
        sc = self.scene
        eps = self.epsilon if epsilon is None else epsilon
        turns = torch.zeros((), dtype=torch.int64, device=self.device)
        sc.begin_tick()
        for k in range(sc.N):
            obs, will = sc.encode(k, L.OBS_CLIP)
            with torch.no_grad():
                q = self.q_values(self.policy_network, obs)
            a_nn = q.argmax(1).to(torch.int8)
            uk = torch.rand(sc.E, device=self.device, generator=self.gen) if u is None else u[k]
            a_in = torch.where(uk < eps, a_nn, torch.full_like(a_nn, -1))
            rew, done, acted, _ = sc.step_turn(k, L.PROTO_INTELLIGENT, a_in)
            a_exec = sc.last_action
            turns += acted.sum()
            self.reward_acc += rew.double().sum()
            if not train:
                continue
            nobs, _ = sc.encode(k, L.OBS_CLIP, post=True)
            with torch.no_grad():
                qn = self.q_values(self.policy_network, nobs)
                # store_transition, MADQN.py:113-128
                storable = (acted > 0) & (a_exec >= 0) & (a_exec < q.shape[1])
                idx = a_exec.clamp(0, q.shape[1] - 1).long().unsqueeze(1)
                old_val = q.gather(1, idx).squeeze(1)
                new_val = rew + self.GAMMA * qn.max(1)[0] * (done == 0)
                prio = ((old_val - new_val).abs().double() + 0.01) ** 0.6
                act_store = torch.where(storable, a_exec, torch.full_like(a_exec, -1))
            self.memory.push(act_store.contiguous(), rew.contiguous(), done.contiguous(), obs.contiguous(),
                             nobs.contiguous(), prio.contiguous())
            self._stored_dev += storable.sum()
            if not self._learning:
                # MADQN.py:131 gates on tree.n_entries: rows the push dropped (no turn, guidance STOP) do not
                # count.  One synchronising read per sub-step until the threshold is passed, none afterwards.
                # (Several ranks pass it together: their update counts must agree for the all-reduce.)
                n = torch.tensor([float(self.memory.info()[0])], device=self.device)
                if self.world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(n, op=dist.ReduceOp.MIN)
                self._learning = n.item() >= self.start_training_info_number
            if self._learning:
                n_upd = self.updates_per_substep
                if self.updates_per_transition is not None:
                    cnt = storable.sum().float().reshape(1)
                    if self.world > 1:  # the ranks' update counts must agree (one all-reduce per update)
                        import torch.distributed as dist
                        dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
                    self._upd_carry += float(cnt.item()) * self.updates_per_transition
                    n_upd = int(self._upd_carry)
                    self._upd_carry -= n_upd
                for _ in range(n_upd):
                    self.update_network()
        self.finish_update()
        sc.end_tick()
        return turns

    # ------------------------------------------------------------------ learning
    @torch.no_grad()
    def _adagrad_step(self):
        """torch.optim.Adagrad's update on the (rank-averaged) gradient held in the flat buffer"""
        grads = self._views
        if self.world > 1:
            torch._foreach_div_(grads, float(self.world))
        torch._foreach_addcmul_(self._sums, grads, grads, value=1.0)
        std = torch._foreach_sqrt(self._sums)
        torch._foreach_add_(std, self.adagrad_eps)
        torch._foreach_addcdiv_(self._params, grads, std, value=-self.lr)

    def _backward_body(self):
        """MADQN.py:134-171 on a device-sampled batch up to the gradient (everything here is capturable:
        kernel launches on the current stream and tensor ops on static buffers)"""
        b = self._sb
        p = lambda t: t.data_ptr()
        L.check(self.memory.lib.agv_replay_sample(self.memory._h, self.batch_size, p(self._u), p(b["slot"]), p(b["prio"]),
                                                  p(b["obs"]), p(b["act"]), p(b["reward"]), p(b["next_obs"]),
                                                  p(b["done"]), self.memory._stream()), "agv_replay_sample")
        # (without autocast the bf16 0/1 states are widened to fp32 -- exact)
        cl = lambda x: (x if self.autocast else x.float()).contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            loss = ddqn_loss(lambda x: self._train_net(cl(x)).float(), lambda x: self.target_network(cl(x)).float(),
                             b["obs"], b["act"].long().clamp(min=0).unsqueeze(1), b["reward"].unsqueeze(1),
                             b["next_obs"], (b["done"] != 0).float().unsqueeze(1), self.GAMMA, self.loss_function)
        loss.backward()
        with torch.no_grad():
            torch._foreach_copy_(self._views, [q.grad for q in self._params])
        self._loss.copy_(loss.detach())

    def _learn_once(self):
        """gradient -> [all-reduce] -> Adagrad, eagerly"""
        for q in self._params:
            q.grad = None
        self._backward_body()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._flat)
        self._adagrad_step()

    def finish_update(self):
        """several ranks: wait for the gradient all-reduce in flight and apply its Adagrad half"""
        if self._pending is None:
            return
        torch.cuda.current_stream(self.device).wait_event(self._pending)
        self._pending = None
        if self._graph_b is not None:
            self._graph_b.replay()
        else:
            self._adagrad_step()

    def _all_reduce_then_step(self):
        """gradient in self._flat -> mean over ranks -> Adagrad; with overlap the step is deferred"""
        import torch.distributed as dist
        if not self.overlap_allreduce:
            dist.all_reduce(self._flat)
            if self._graph_b is not None:
                self._graph_b.replay()
            else:
                self._adagrad_step()
            return
        cur = torch.cuda.current_stream(self.device)
        self._comm.wait_stream(cur)
        with torch.cuda.stream(self._comm):
            dist.all_reduce(self._flat)
            ev = torch.cuda.Event()
            ev.record(self._comm)
        self._pending = ev

    def update_network(self):
        """one optimiser step; the first three run eagerly (warm-up on a side stream), then the step is
        captured once and replayed"""
        self.finish_update()
        if self.learn_step_counter % self.TARGET_REPLACE_ITER == 0:
            self.target_network.load_state_dict(self.policy_network.state_dict())
        self.learn_step_counter += 1
        self._u.copy_(torch.rand(self.batch_size, dtype=torch.float64, device=self.device, generator=self.gen))
        if self._graph is not None:
            self._graph.replay()
            if self.world > 1:
                self._all_reduce_then_step()
        elif not self.graph_learner:
            for q in self._params:
                q.grad = None
            self._backward_body()
            if self.world > 1:
                self._all_reduce_then_step()
            else:
                self._adagrad_step()
        elif self._warm < 3:
            self._warm += 1
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._learn_once()
            cur.wait_stream(side)
        else:
            for q in self._params:
                q.grad = None
            g = torch.cuda.CUDAGraph()
            if self.world == 1:
                with torch.cuda.graph(g):
                    self._backward_body()
                    self._adagrad_step()
                self._graph = g
                g.replay()  # the capture itself does not execute the step
            else:
                gb = torch.cuda.CUDAGraph()
                # thread-local capture mode: the NCCL watchdog thread polls events meanwhile
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._backward_body()
                with torch.cuda.graph(gb, capture_error_mode="thread_local"):
                    self._adagrad_step()
                self._graph, self._graph_b = g, gb
                g.replay()
                self._all_reduce_then_step()
        self._loss_ring[(self.learn_step_counter - 1) % self.LOSS_RING].copy_(self._loss)
        return self._loss

    def change_explore_rate(self, times):
        """MADQN.py:182-190"""
        if self.epsilon_count >= times:
            self.epsilon = self.epsilon_end
        else:
            self.epsilon = self.epsilon + (self.epsilon_end - self.epsilon_start) / times
        self.epsilon_count += 1

    def save(self, path_policy, path_target, path_state=None):
        """policy_net.pt / target_net.pt as whole-module pickles under the REFERENCE's class path
        (Controller.py:161-172), so that the reference's `use_NN` torch.load reads them
        (algorithm/checkpoint.py); `path_state` additionally keeps what a resumed training run needs and the
        reference never saves: the Adagrad accumulators, the update counter, epsilon."""
        from ..checkpoint import save_module
        self.finish_update()
        save_module(self.policy_network, path_policy)
        save_module(self.target_network, path_target)
        if path_state is not None:
            torch.save({"adagrad_sum": [t.detach().cpu() for t in self._sums],
                        "learn_step_counter": self.learn_step_counter, "epsilon": self.epsilon,
                        "epsilon_count": self.epsilon_count}, path_state)

    def load(self, path_policy, path_target=None, path_state=None):
        """reads checkpoints written by save() or by the reference (whole-module pickles or state_dicts)"""
        from ..checkpoint import load_module
        self.finish_update()

        def state_of(path):
            obj = load_module(path, map_location="cpu")
            return obj.state_dict() if isinstance(obj, nn.Module) else obj

        self.policy_network.load_state_dict(state_of(path_policy))
        self.target_network.load_state_dict(state_of(path_target) if path_target else self.policy_network.state_dict())
        if path_state is not None:
            st = torch.load(path_state, map_location="cpu")
            with torch.no_grad():
                for dst, src in zip(self._sums, st["adagrad_sum"]):
                    dst.copy_(src)
            self.learn_step_counter = int(st["learn_step_counter"])
            self.epsilon, self.epsilon_count = float(st["epsilon"]), int(st["epsilon_count"])
