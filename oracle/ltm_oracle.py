"""TEST INFRASTRUCTURE — restatement of the reference's long-term memory (memory/ltm.py:8-188) with pandas, on string
keys (any hashable stands in for pymatgen's `composition.reduced_formula` / element tuple): the same DataFrame
operations, line for line, minus the pymatgen calls that produce the keys.  Only tests may import this."""
import numpy as np
import pandas as pd


class LongTimeMemOracle:
    def __init__(self):
        self.memory = pd.DataFrame(columns=["comp", "ele_comb", "reward", "RL_step"])
        self.unique_comps = []

    def extend(self, comps, ele_comb, rewards, step):                     # ltm.py:30-63
        df = pd.DataFrame.from_dict({"comp": list(comps), "ele_comb": list(ele_comb), "reward": np.asarray(rewards, float),
                                     "RL_step": [step] * len(comps)})
        self.memory = pd.concat([self.memory, df]) if len(self.memory) > 0 else df
        self.unique_comps = self.memory["comp"].unique()

    def div_filter(self, values, rewards, tol=10, buff=20, method="composition"):      # ltm.py:65-109
        assert tol < buff
        key = "comp" if method == "composition" else "ele_comb"
        new_rewards, penalty_idx, tol_n, buff_n = [], [], 0, 0
        for i, v in enumerate(values):
            occ = self.memory[key].value_counts().get(v, 0)
            if occ <= tol:
                new_rewards.append(rewards[i])
            elif occ > tol and occ < buff:
                new_rewards.append(rewards[i] * (buff - occ) / (buff - tol))
                tol_n += 1
            else:
                new_rewards.append(0.0)
                penalty_idx.append(i)
                buff_n += 1
        return np.array(new_rewards), penalty_idx, tol_n, buff_n

    def calc_metrics(self, thred, budget=3000, num_candidate=100):        # ltm.py:111-133
        _df = self.memory.sort_values("reward", ascending=False)
        unique_df = _df.drop_duplicates(subset=["comp"])
        candidates = (unique_df["reward"] > thred).sum()
        calc_cost = len(self.memory)
        burden = calc_cost / candidates if candidates >= num_candidate else None
        div_ratio = len(self.unique_comps) / calc_cost if calc_cost <= budget else None
        return burden, div_ratio

    def get_baseline(self, step, prev=3):                                 # ltm.py:135-137
        return self.memory[self.memory["RL_step"] > step - prev]["reward"].mean()
