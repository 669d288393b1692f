"""Interface a variational autoencoder must comply with (mirrors mvae/vae.py:7-81)."""
import abc


class VAE(abc.ABC):
    @abc.abstractmethod
    def sample(self):
        """From z_dim input produce an input_dim output"""
        raise NotImplementedError()

    @abc.abstractmethod
    def predict(self):
        """From input_dim input produce an input_dim output"""
        raise NotImplementedError()

    @abc.abstractmethod
    def encode(self):
        """From input_dim input produce an z_dim output"""
        raise NotImplementedError()

    @property
    @abc.abstractmethod
    def z_dim(self) -> int:
        raise NotImplementedError()

    @property
    @abc.abstractmethod
    def input_dim(self):
        raise NotImplementedError()

    @property
    @abc.abstractmethod
    def model_decode(self):
        raise NotImplementedError()

    @property
    @abc.abstractmethod
    def model_encode(self):
        raise NotImplementedError()

    @property
    @abc.abstractmethod
    def model_trainable(self):
        raise NotImplementedError()
