import torch


def to_categorical(y, num_classes):
    new_y = torch.eye(num_classes)[y.cpu().data.numpy(),]
    return new_y.cuda() if y.is_cuda else new_y
