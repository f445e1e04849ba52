"""pipelines/pipeline.py:7-202 on the B200 engine: the same orchestration object (train / test /
publish / save / load / save_trajectory, archive and report directories, metadata.json) driving
Rollout_Buffer.sample() -> fused rollout kernel and Algorithm.learn() -> advantage + update kernels.

File formats are the reference's, so `reports/**` checkpoints written by either side load in the other:
    policy.pt       torch state dict (`network.{0,2,..}.{weight,bias}`; {'actor':..,'critic':..} for ActorCritic)
    optimizer.pth   GRPO (grpo.py:150-160)   /   optimizer.pt   PPO (ppo.py:207-225)
    reward.csv      one mean episodic return per epoch (rollout_buffer.py:115-126)
    metadata.json   pipeline.py:120-139
The visualizer and publisher are optional collaborators (matplotlib is out of scope here): any object
with the reference's `initialize/plot/render/metadata` resp. `publish/report/metadata` methods plugs in.
"""
from __future__ import annotations

import datetime
import json
import os
from typing import Any, Callable, Dict, Optional


class Pipeline:
    def __init__(self, test_name: str, checkpoint_name: str, env_fn: Callable[[], Any], policy: Any, algorithm: Any,
                 rollout_manager: Any, buffer: Any, visualizer: Optional[Any] = None, publisher: Optional[Any] = None,
                 logger: Optional[Any] = None, load_path: Optional[str] = None, save_freq: int = 10,
                 render_freq: int = 40, root: str = ".") -> None:
        self.test_name, self.checkpoint_name = test_name, checkpoint_name
        self.env_fn = env_fn
        self.env = env_fn()
        self.env_name = self.env.env_name
        self.policy, self.algorithm = policy, algorithm
        self.rollout_manager, self.buffer = rollout_manager, buffer
        self.visualizer, self.publisher, self.logger = visualizer, publisher, logger
        self.load_path, self.save_freq, self.render_freq = load_path, save_freq, render_freq
        self.root = root
        self.today = datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")
        self.loaded_metadata: Optional[Dict[str, Any]] = None
        if load_path is not None:
            self.load()
        self.initialize()

    def initialize(self) -> None:
        """pipeline.py:75-91."""
        self.archive_path = os.path.join(self.root, "archive", self.env_name, self.test_name, self.checkpoint_name)
        self.publish_path = os.path.join(self.root, "reports", self.env_name, self.test_name, self.checkpoint_name)
        os.makedirs(self.archive_path, exist_ok=True)
        if self.load_path is not None:
            self.loaded_metadata = self.load_metadata(os.path.join(self.load_path, "metadata.json"))
        metadata = self.get_metadata()
        if self.visualizer is not None:
            self.visualizer.initialize(metadata)

    def load(self) -> None:
        """pipeline.py:93-102."""
        if self.load_path is not None:
            self.algorithm.load(self.load_path)
            self.policy.load(self.load_path)
            self.buffer.load(self.load_path)

    def save(self, path: str) -> None:
        """pipeline.py:104-118."""
        os.makedirs(path, exist_ok=True)
        self.algorithm.save(path)
        self.policy.save(path)
        self.buffer.save(path)
        with open(os.path.join(path, "metadata.json"), "w") as f:
            json.dump(self.get_metadata(), f, indent=4)

    def get_metadata(self) -> Dict[str, Any]:
        """pipeline.py:120-139."""
        return {
            "test_name": self.test_name,
            "checkpoint_name": self.checkpoint_name,
            "creation_date": self.today,
            "env_name": self.env_name,
            "policy": self.policy.metadata(),
            "algorithm": self.algorithm.metadata(),
            "buffer": self.buffer.metadata(),
            "visualizer": self.visualizer.metadata() if self.visualizer is not None else {},
            "publisher": self.publisher.metadata() if self.publisher is not None else {},
            "logger": self.logger.metadata() if self.logger is not None else {},
        }

    def load_metadata(self, path: str) -> Dict[str, Any]:
        """pipeline.py:141-153."""
        with open(path, "r") as f:
            return json.load(f)

    def train(self, epochs: int) -> None:
        """pipeline.py:155-176: sample -> learn, plot/render/save at their frequencies."""
        for epoch in range(epochs):
            self.buffer.sample()
            self.algorithm.learn(self.buffer)
            if hasattr(self.visualizer, "plot"):
                self.visualizer.plot()
            if self.visualizer is not None and epoch % self.render_freq == 0:
                self.visualizer.render()
            if epoch % self.save_freq == 0:
                self.save(self.archive_path)

    def test(self) -> None:
        """pipeline.py:178-182."""
        self.buffer.sample()

    def publish(self) -> None:
        """pipeline.py:184-193."""
        os.makedirs(self.publish_path, exist_ok=True)
        self.buffer.sample()
        if self.publisher is not None:
            self.publisher.publish(self.publish_path)
            self.publisher.report(self.publish_path, self.get_metadata())
        self.save(self.publish_path)

    def save_trajectory(self) -> None:
        """pipeline.py:195-201."""
        self.buffer.sample()
        self.buffer.save_trajectory(self.archive_path)
