"""Probe: does CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) work between two processes in this container?
The cross-process forward hop (a6/X7) maps the ingress rank's request window into the owner rank's address space."""
import torch
import torch.multiprocessing as mp


def child(q, r):
    t = q.get()
    t += 1
    torch.cuda.synchronize()
    r.put("ok")


if __name__ == "__main__":
    mp.set_start_method("spawn")
    q, r = mp.Queue(), mp.Queue()
    p = mp.Process(target=child, args=(q, r))
    p.start()
    x = torch.zeros(1024, device="cuda")
    q.put(x)
    print("child:", r.get(timeout=120))
    torch.cuda.synchronize()
    print("IPC_OK" if float(x.sum()) == 1024 else "IPC_BAD", float(x.sum()))
    p.join()
